mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
for B in 16 32; do python bench.py --steps 6 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/bench_b$B.log 2>&1; tail -1 gpurun_out/bench_b$B.log | cut -c1-170; done
python bench.py --model retinanet --steps 8 --warmup 3 --batch 16 --no-cpu-baseline --layers gpurun_out/layers_retina2.tsv > gpurun_out/bench_retina2.log 2>&1; tail -1 gpurun_out/bench_retina2.log | cut -c1-170
