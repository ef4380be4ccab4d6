#!/bin/bash
# tools/ab.sh [render-levels] — run tools/probe_scale.py for librt_core.so and every build/rt_*.so variant (A/B tuning runs).
LEVELS=${1:-6,8}
fmt() { python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{\"probe\": \"render'):
        d=json.loads(l); print('  %8d tris  %.3f ms/frame  all %5.0f  '%(d['triangles'], d['ms_per_frame'], d['mrays_per_s_all']) + '  '.join('%s %5.0f'%(k,d[k]['mrays_per_s']) for k in ('primary','secondary','shadow')) + '  img ' + d.get('image_sha1',''))
    elif l.startswith('{\"probe\": \"build'):
        d=json.loads(l); print('  build %9d tris %.3f ms  %.0f Mtri/s'%(d['triangles'], d['ms'], d['mtri_per_s']))
    else: print(l[:300])
"; }
echo "== default"; python tools/probe_scale.py --build "${BUILD:-}" --render $LEVELS 2>&1 | tail -${TAIL:-4} | fmt
for so in build/rt_*.so; do [ -e "$so" ] || continue; echo "== $so"; RT_CORE_LIB=$PWD/$so python tools/probe_scale.py --build "${BUILD:-}" --render $LEVELS 2>&1 | tail -${TAIL:-4} | fmt; done
