# A/B timing of library variants on the same box: tools/ab.sh "<quick_bench args>" name1 name2 ...  ("cur" = in-tree build)
args=$1; shift
for rep in 1 2; do
for v in "$@"; do
  lib=$PWD/coupe_b200/lib/variants/lib$v.so
  [ "$v" = cur ] && lib=$PWD/coupe_b200/lib/libcoupe_b200.so
  echo "== $v"
  COUPE_B200_LIB=$lib python tools/quick_bench.py $args --reps 5 2>&1 | grep -E "best"
done
done
