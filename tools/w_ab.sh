B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e"
for w in 0 888 592 296; do
  SEPCORE_STRIP_WARPS=$w $B --size 512 --shift 128 --sources 3 --seconds 8 --window hann > gpurun_out/w9_512_$w.log 2>&1
  SEPCORE_STRIP_WARPS=$w $B --sources 1 > gpurun_out/w9_c1_$w.log 2>&1
done
