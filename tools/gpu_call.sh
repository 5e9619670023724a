cd "${GRAFT_REPO_ROOT:-.}"
for v in base sh5 sh6 an5 fi5 all5 an1 fi1 sh1; do
  B2PT_LIB=build/variants/libb2pt_$v.so timeout 300 python tools/exp_walk.py $v 2>&1 | tail -n 3 | cut -c1-215
done
