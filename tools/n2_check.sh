mkdir -p gpurun_out
python -m pytest tests/test_tiled_nccl.py -x -q 2>&1 | tail -5 > gpurun_out/t_nccl.log; cat gpurun_out/t_nccl.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/tiled_run.py 16384 16384 4 1 5 > gpurun_out/tiled_16384_n2.txt 2>&1; tail -8 gpurun_out/tiled_16384_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/tiled_run.py 32768 32768 4 0 5 > gpurun_out/tiled_32768_n2.txt 2>&1; tail -6 gpurun_out/tiled_32768_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -c 400 gpurun_out/bench_n2.err
