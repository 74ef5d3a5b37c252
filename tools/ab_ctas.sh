#!/bin/bash
# A/B of the resident footprint-kernel CTAs per SM x pipeline form x library build on the default bench.
# usage: CTAS="1 2 3" PIPES="off deep" LIBS="- path/to/other.so" bash tools/ab_ctas.sh
for lib in ${LIBS:--}; do for pipe in ${PIPES:-deep off}; do for c in ${CTAS:-2 3}; do
  if [ "$lib" = "-" ]; then unset LP_B200_LIB; else export LP_B200_LIB=$lib; fi
  LP_RASTER_CTAS=$c python bench.py --steps 200 --warmup 10 --no-e2e --no-strong --cpu-views 0 --pipeline $pipe 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $pipe ctas $c', round(1e3*d['ms_per_step'],1), {k:round(v,1) for k,v in d['roofline']['kernels_us'].items()})"
done; done; done
