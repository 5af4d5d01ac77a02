python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pmavg or computePMparams or multi_neuron or plan_relaunch or sharded_halves or dense" > gpurun_out/pytest_r02b.log 2>&1; tail -25 gpurun_out/pytest_r02b.log
python tools/gpu_offnode.py > gpurun_out/offnode_r02b.json 2> gpurun_out/offnode_r02b.err; cat gpurun_out/offnode_r02b.json; tail -3 gpurun_out/offnode_r02b.err
