#!/bin/bash
# round-2 trip A: new parity tests (C2 full size, reference heads, masks, C5 shapes) + graph probe
mkdir -p gpurun_out
python -m pytest tests/test_c2_parity_gpu.py tests/test_reference_heads_gpu.py -x -q -m gpu -s 2>&1 | tail -40 > gpurun_out/r2a_tests.log
python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -15 >> gpurun_out/r2a_tests.log
timeout 300 python tools/gpu_graph_probe.py > gpurun_out/r2a_graph_probe.json 2> gpurun_out/r2a_graph_probe.err
timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err
cat gpurun_out/r2a_tests.log | tail -30; cat gpurun_out/r2a_graph_probe.json; tail -3 gpurun_out/r2a_graph_probe.err; cat gpurun_out/r2a_bench_reference.json
