#!/bin/bash
# A/B prebuilt libraries on the GPU box: usage tools/ab.sh [-a "bench args"] libA.so libB.so ...   (restores the in-tree library at the end)
extra=""
if [ "$1" = "-a" ]; then extra="$2"; shift 2; fi
cp radex_emcee_b200/libradex_b200.so /tmp/lib_keep.so
for lib in "$@"; do
  cp "$lib" radex_emcee_b200/libradex_b200.so
  python bench.py --log2n 17 --steps 2 --warmup 3 --no-cpu $extra 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$extra', 'solves/s %.4g iters/s %.4g frac %.3f it/solve %.1f cached %.3f capt %.3f inval %.3f maxit %.3f'%(d['value'], d['matrix_iterations_per_s'], d['roofline']['frac'], d['iters_per_solve'], d['frac_iterations_cached'], d['captures_per_solve'], d['invalidations_per_solve'], d['frac_at_maxiter']))"
done
cp /tmp/lib_keep.so radex_emcee_b200/libradex_b200.so
