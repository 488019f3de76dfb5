# Development aid: tools/fused_bench.py over library variants built by tools/dev_variants.sh.
#   tools/variant_sweep.sh "<tag>[:ENV=VAL[,ENV=VAL]]" ...
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [ "$spec" != "$v" ]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  echo "variant=$spec"
  env $envs APOLLO_B200_LIB=build/libmsda_$v.so timeout 200 python tools/fused_bench.py 2>&1 | tail -1
done
