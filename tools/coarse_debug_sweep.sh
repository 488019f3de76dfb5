# Development aid: device time of sca_coarse_scatter with parts switched off (MSDA_COARSE_DEBUG bits:
# 1 = no build / erase, 2 = no MMAs, 4 = no record fetch), from an ncu launch list.
for d in ${@:-0 1 2 7}; do
echo "debug=$d"
MSDA_COARSE_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"coarse_scatter|fused_bwd" -s 4 -c 4 --csv python tools/fused_bench.py --which sca --iters 2 2>/dev/null | grep -E "coarse_scatter|fused_bwd" | awk -F, '{print substr($5,1,30), $NF}'
done
