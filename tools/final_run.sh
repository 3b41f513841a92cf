set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02g_tests.log
tail -3 gpurun_out/r02g_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1; tail -2 gpurun_out/r02g_smoke.log
timeout 1200 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -1 gpurun_out/r02g_bench.json | cut -c1-400
