#!/bin/bash
# round-end verification on the GPU box: full GPU parity suite, smoke(), default bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/verify_tests.txt 2>&1
tail -4 gpurun_out/verify_tests.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err
tail -c 3000 gpurun_out/verify_bench.json
