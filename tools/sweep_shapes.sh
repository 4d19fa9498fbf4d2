#!/bin/bash
# tuning sweep: build variant (tools/build_variants.py name=flags ...) x forced launch shape (BF_REFINE_SHAPE), C1 / C4 kernel-level cases
# usage: bash tools/sweep_shapes.sh name1 name2 ...   (names of lib_<name>.so under boxfusion_b200/lib/variants)
for v in "$@"; do
  for shape in auto 1,256 2,256 4,256 8,256 2,128 4,128 8,128; do
    if [ "$shape" = auto ]; then unset BF_REFINE_SHAPE; else export BF_REFINE_SHAPE=$shape; fi
    BOXFUSION_B200_LIB=$PWD/boxfusion_b200/lib/variants/lib_$v.so python tools/kernel_bench.py --cases c1,c4f 2>&1 | python -c "
import sys,json
o=[]
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    if 'case' in d: o.append('%s %.3f' % (d['case'][:9], d['ms']))
print('$v', '$shape', ' | '.join(o))
"
  done
done
