#!/bin/bash
# usage: collect_profiles.sh <TAG>  — copies the round evidence from gpurun_out/ into profiles/ and regenerates the summaries
TAG=$1
for c in M P B T reference; do cp gpurun_out/${TAG}_bench_$c.json profiles/r1_bench_$c.json; done
cp gpurun_out/${TAG}_pytest_gpu.txt profiles/r1_pytest_gpu.txt
cp gpurun_out/${TAG}_launches_M.csv profiles/r1_launches_M.csv
[ -f gpurun_out/${TAG}_phases_M.txt ] && cp gpurun_out/${TAG}_phases_M.txt profiles/r1_phases_M.txt
CMD="ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 python bench.py --config %s --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-separation"
for c in M P; do
python scripts/make_profile_md.py gpurun_out/${TAG}_prof_$c.ncu-rep profiles/r1_assemble_${c}_final.md "round 1, final capture (session 4) — assemble_kernel, config $c" "$(printf "$CMD" $c)" > /dev/null
echo "" >> profiles/r1_assemble_${c}_final.md; echo "## shared-memory wavefronts by SASS cluster (scripts/ncu_smem_breakdown.py)" >> profiles/r1_assemble_${c}_final.md; echo "" >> profiles/r1_assemble_${c}_final.md
python scripts/ncu_smem_breakdown.py gpurun_out/${TAG}_prof_$c.ncu-rep >> profiles/r1_assemble_${c}_final.md
done
python scripts/make_profile_md.py gpurun_out/${TAG}_prof_pcg_M.ncu-rep profiles/r1_pcg_M.md "round 1 (session 4) — one Jacobi-PCG iteration on K_aa of config M (24.0 M rows, 504 M stored entries)" "ncu --set full --clock-control none -k regex:spmv_dot_kernel|update_kernel|direction_kernel -s 6 -c 3 python bench.py --config M --steps 1 --warmup 1 --no-e2e --no-cpu-baseline" > /dev/null
python - <<'PY'
import csv, json, re
rows=[r for r in csv.reader(open('profiles/r1_launches_M.csv')) if len(r)>5]
hdr=rows[0]; data=rows[1:]
ix={h:i for i,h in enumerate(hdr)}
ks=[(r[ix['ID']], r[ix['Kernel Name']], r[ix['Grid Size']], r[ix['Block Size']], float(r[ix['Metric Value']])/1e3) for r in data if r[ix['Metric Name']]=='gpu__time_duration.sum']
last=ks[-4:]
tot=sum(k[4] for k in last)
def short(n):
    m=re.search(r'(\w+_kernel<[^>]*>|\w+_kernel)', n); return m.group(1) if m else n
bm=json.loads(open('profiles/r1_bench_M.json').read().strip().splitlines()[-1])
r=bm['roofline']
L=["# round 1 — ncu launch list, default bench workload (config M, 10M mixed elements)","",
"`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-separation` (full list: `r1_launches_M.csv`, %d launches: load/validate, symbolic pass, then the numeric passes)"%len(ks),"",
"One numeric pass (= one bench step) is four launches; the three element-record kernels run on three streams (side by side in a live run, serialised under ncu):","",
"| # | kernel | grid x block | time (us) | share of step |","|---|---|---|---|---|"]
for k in last:
    L.append("| %s | %s | %s x %s | %.1f | %.1f %% |"%(k[0],short(k[1]),k[2],k[3],k[4],100*k[4]/tot))
asm=[k for k in last if 'assemble' in k[1]][0]
L+=["","Step total under ncu (serialised, cold cache): %.3f ms; `assemble_kernel` share %.1f %%. Live CUDA-event timing in `r1_bench_M.json`: step %.3f ms, assemble %.3f ms (%.1f %%), element records %.3f ms (overlapped: less than the %.3f ms sum of the three serialised launches) — the shares agree."%(tot/1e3,100*asm[4]/tot,bm['ms_per_step'],r['kernel_ms'],100*r['kernel_ms']/bm['ms_per_step'],r['prep_ms'],(tot-asm[4])/1e3)]
open('profiles/r1_launches_M.md','w').write("\n".join(L)+"\n")
def dram(md):
    t=open(md).read()
    rd=float(re.search(r"dram__bytes_read.sum \| Gbyte \| ([\d.]+)",t).group(1)); wr=float(re.search(r"dram__bytes_write.sum \| Gbyte \| ([\d.]+)",t).group(1))
    return int((rd+wr)*1e9)
tj={"M":{"dram_bytes_per_launch":dram('profiles/r1_assemble_M_final.md'),"kernel":"assemble_kernel<64, 1, 1>","source":"profiles/r1_assemble_M_final.md"},
    "P":{"dram_bytes_per_launch":dram('profiles/r1_assemble_P_final.md'),"kernel":"assemble_kernel<32, 0, 1>","source":"profiles/r1_assemble_P_final.md"}}
json.dump(tj,open('profiles/traffic.json','w'),indent=1)
for c in 'MPBT':
    d=json.loads(open('profiles/r1_bench_%s.json'%c).read().strip().splitlines()[-1]); r=d['roofline']; e=d['e2e']; s=d['separation']
    print(c, d['config']['elements'], round(d['ms_per_step'],3), round(d['value']/1e9,3), round(r['kernel_ms'],3), round(r['achieved']), round(100*r['frac'],1), round(e['value']/1e6,2), round(e.get('seconds_per_step',0),3), round(s['ms'],2))
PY
