B="python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e"
for v in "--shift 64 --window hann" "--sources 1" "--sources 1 --shift 64"; do
  tag=$(echo "$v" | tr -d ' -')
  SEPCORE_WSTRIP_WARPS=592 $B $v > gpurun_out/w7_dual592_$tag.log 2>&1
  SEPCORE_WSTRIP_WARPS=296 $B $v > gpurun_out/w7_dual296_$tag.log 2>&1
  SEPCORE_FORCE_HALFWARP=1 $B $v > gpurun_out/w7_old_$tag.log 2>&1
done
SEPCORE_WSTRIP_WARPS=444 $B > gpurun_out/w7_dual444_.log 2>&1
SEPCORE_WSTRIP_WARPS=296 $B > gpurun_out/w7_dual296_.log 2>&1
SEPCORE_WSTRIP_WARPS=592 $B --batch 512 > gpurun_out/w7_dual592_b512.log 2>&1
