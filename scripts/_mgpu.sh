python -m pytest tests -m gpu -x -q -k "multi_device" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/bench_n2_r02.json 2> gpurun_out/bench_n2_r02.err; echo "rc=$?"; cat gpurun_out/bench_n2_r02.json; tail -5 gpurun_out/bench_n2_r02.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-budget 5 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 -m pysonic_b200.run_lookups -n RS -a 32 -f 500 -A 0 50 300 --mpi -o gpurun_out/cli_n2 -y 2>&1 | tail -3; ls -la gpurun_out/cli_n2
