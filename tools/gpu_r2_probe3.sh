#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_text_gpu.py tests/test_k1_gpu.py -m gpu -q -x > gpurun_out/p3_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/p3_pytest.log
for lib in "" $K6_VARIANTS; do
  R3D_LIB_PATH=$PWD/3d_reconstruction_system_b200/libr3d_b200$lib.so timeout 300 python tools/k6_probe.py 64 $PROBE_ARGS > gpurun_out/k6_probe$lib.json 2> gpurun_out/k6_probe$lib.err || tail -3 gpurun_out/k6_probe$lib.err
  python - "$lib" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/k6_probe%s.json'%sys.argv[1]))
    t=d['text_rows']
    c=d.get('compact_mode') or {}
    print('%-7s'%(sys.argv[1] or 'deflt'),{k:(round(v['ms'],3),round(v['points_per_s']/1e9,2),round(v['frac_of_hbm_peak'],3),v.get('parity_ok')) for k,v in t.items() if isinstance(v,dict)}, 'compact ms',c.get('ms'),'frac',c.get('frac_of_hbm_peak'))
except Exception as e:
    print(sys.argv[1],'failed',e)
PY
done
