mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"
tail -c 1500 gpurun_out/bench_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_n2_ref.log 2>&1; echo "n2 ref rc=$?"
tail -c 600 gpurun_out/bench_n2_ref.log
