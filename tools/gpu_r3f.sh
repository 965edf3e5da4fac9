#!/bin/bash
# round 2, call 3F (1 GPU): where does c4's direct kernel spend its 20 us? ncu --set full, summary + per-opcode / per-region stall samples
mkdir -p gpurun_out /tmp/rep
ncu --set full --clock-control none --import-source on -k "regex:direct_tile" -s 3 -c 1 -o /tmp/rep/c4 -f python tools/run_c4.py > gpurun_out/ncu_c4_r02d.log 2>&1
python tools/summarize_ncu.py /tmp/rep/c4.ncu-rep gpurun_out/r02d_ncu_full_c4 | cut -c1-900
ncu -i /tmp/rep/c4.ncu-rep --page source --csv 2>/dev/null > /tmp/rep/c4_src.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/rep/c4_src.csv')))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
buck=[]; tot=0
for k,r in enumerate(rows[2:]):
    if len(r)<len(hdr): continue
    s=int(r[ix['# Samples']] or 0); tot+=s
    b=k//40
    while len(buck)<=b: buck.append({'s':0,'ops':{}, 'st':{}})
    buck[b]['s']+=s
    src=r[ix['Source']].strip().split()
    op=src[1] if src[0].startswith('@') else src[0]
    buck[b]['ops'][op]=buck[b]['ops'].get(op,0)+1
    for h in stalls:
        v=int(r[ix[h]] or 0)
        if v: buck[b]['st'][h]=buck[b]['st'].get(h,0)+v
out=['total samples %d' % tot]
for b,d in enumerate(buck):
    ops=sorted(d['ops'].items(), key=lambda x:-x[1])[:4]
    st=sorted(d['st'].items(), key=lambda x:-x[1])[:4]
    out.append('%d %d %s | %s' % (b*40, d['s'], ' '.join(f'{o}:{c}' for o,c in ops), ' '.join(f'{o[6:]}:{c}' for o,c in st)))
open('gpurun_out/r02d_c4_stall_regions.txt','w').write('\n'.join(out)+'\n')
print('\n'.join(out))
PY
