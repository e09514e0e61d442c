python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/w12_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w12_smoke.log 2>&1
B="python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-e2e"
$B > gpurun_out/w12_cfg2.log 2>&1
SEPCORE_SINGLE_LAUNCH=0 $B > gpurun_out/w12_cfg2_multi.log 2>&1
$B --shift 64 --window hann > gpurun_out/w12_hop64.log 2>&1
SEPCORE_SINGLE_LAUNCH=0 $B --shift 64 --window hann > gpurun_out/w12_hop64_multi.log 2>&1
$B --sources 1 > gpurun_out/w12_c1.log 2>&1
SEPCORE_SINGLE_LAUNCH=0 $B --sources 1 > gpurun_out/w12_c1_multi.log 2>&1
$B --steps 100 --size 512 --shift 128 --sources 3 --seconds 8 --window hann > gpurun_out/w12_cfg4.log 2>&1
SEPCORE_SINGLE_LAUNCH=0 $B --steps 100 --size 512 --shift 128 --sources 3 --seconds 8 --window hann > gpurun_out/w12_cfg4_multi.log 2>&1
$B --streams 8 > gpurun_out/w12_cfg2_s8.log 2>&1
$B --streams 4 > gpurun_out/w12_cfg2_s4.log 2>&1
