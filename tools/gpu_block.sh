#!/bin/bash
# one MixerBlock fwd+bwd (B/16 shapes, batch 256): in-loop time, then the per-launch ncu duration list
mkdir -p gpurun_out
python tools/one_block.py 20 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gemm_bf16|layernorm|colsum|rowsum|pad_rows|cast_f32" --launch-skip 21 -c 21 --csv --log-file gpurun_out/block_launches.csv python tools/one_block.py 1 > /dev/null 2>&1
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/block_launches.csv')))
h=None;tot=0
for r in rows:
    if len(r)>5 and r[0]=='ID': h=r; continue
    if h and len(r)==len(h):
        d=dict(zip(h,r)); v=float(d['Metric Value'].replace(',',''))/1e3; tot+=v
        print(f"{v:8.1f} us  {re.sub(r'^void |vmlp::','',d['Kernel Name'])[:48]} {d['Grid Size']}")
print(f"sum {tot:.1f} us")
PY
