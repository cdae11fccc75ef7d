#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -x --timeout 200 -p no:cacheprovider > gpurun_out/cE_pytest.log 2>&1
echo "pytest train rc=$?"; tail -n 6 gpurun_out/cE_pytest.log
timeout 400 python -m pytest tests/test_variants_gpu.py -q -k "TRAIN_TC" --timeout 200 -p no:cacheprovider > gpurun_out/cE_variants.log 2>&1
echo "variants rc=$?"; tail -n 4 gpurun_out/cE_variants.log
for tc in 3 5; do echo "TRAIN_TC=$tc"; ROKO_B200_TRAIN_TC=$tc timeout 200 python scripts/train_profile.py 128 20; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cE_train_launches.csv python scripts/train_profile.py 128 2 > gpurun_out/cE_train_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/cE_train_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
L=[(r[ki], float(r[vi].replace(',',''))*(1e-3 if r[ui]=='ns' else 1)) for r in rows[1:]]
# last step: from the last embed_drop_kernel on
st=max(i for i,(k,_) in enumerate(L) if k.startswith('embed_drop'))
agg=collections.OrderedDict()
for k,v in L[st:]:
    k=k[:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda t:-t[1][1]): print(f"{a[1]:9.1f} us {100*a[1]/tot:5.1f}% {a[0]:3d}x {k}")
print("total", round(tot,1), "us over", sum(a[0] for a in agg.values()), "launches")
PY
