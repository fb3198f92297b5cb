#!/bin/bash
# gpurun_out/<tag> (tools/gpu_round.sh) -> profiles/: ncu summaries, bench lines, launch list, ncu_traffic.json
O=gpurun_out/${1:-r01d}; R=${2:-r01}
python tools/ncu_summary.py $O/prof.ncu-rep profiles/${R}_ncu_f32_dif2_512.json $O/launches.csv
python tools/ncu_summary.py $O/prof__dtypef64.ncu-rep profiles/${R}_ncu_f64_dif2_512.json
python tools/ncu_summary.py $O/prof__dif_order0.ncu-rep profiles/${R}_ncu_f32_512.json
python tools/ncu_summary.py $O/prof__dif_order0__dtypef64.ncu-rep profiles/${R}_ncu_f64_512.json
python tools/ncu_summary.py $O/prof__update_type3.ncu-rep profiles/${R}_ncu_f32_iiso_dif2_512.json
python tools/ncu_summary.py $O/prof__update_type3__dif_order0.ncu-rep profiles/${R}_ncu_f32_iiso_512.json
cp $O/launches.csv profiles/${R}_launches_f32_dif2_512.csv
for f in bench_f32_dif2 bench_f64_dif2 bench_f32 bench_f64 bench_f32_iiso_dif2; do cp $O/$f.json profiles/${R}_${f}_512.json; done
cp $O/bench_ref_f32.json profiles/${R}_bench_reference_f32_512.json; cp $O/bench_ref_f64.json profiles/${R}_bench_reference_f64_512.json
python - <<PY
import json
R="$R"
def tr(f): 
    d=json.load(open(f))['launches'][0]; return d['dram_bytes_per_launch'], d['kernel'].split('(')[0]
ents=[]
for dt,ut,do,f in [("f32",0,2,"f32_dif2"),("f64",0,2,"f64_dif2"),("f32",0,0,"f32"),("f64",0,0,"f64"),("f32",3,2,"f32_iiso_dif2"),("f32",3,0,"f32_iiso")]:
    b,k=tr(f"profiles/{R}_ncu_{f}_512.json")
    ents.append({"workload":"c2","dtype":dt,"update_type":ut,"dif_order":do,"dram_bytes_per_launch":b,"source":f"profiles/{R}_ncu_{f}_512.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, 512^3 slab, {k})"})
json.dump(ents, open('profiles/ncu_traffic.json','w'), indent=1)
def L(f): return json.loads(open(f'profiles/{f}').read().strip().splitlines()[-1])
for name,f in [('fp32 SRL_FORWARD + DIF 2',f'{R}_bench_f32_dif2_512.json'),('fp64 SRL_FORWARD + DIF 2',f'{R}_bench_f64_dif2_512.json'),('fp32 SRL_FORWARD',f'{R}_bench_f32_512.json'),('fp64 SRL_FORWARD',f'{R}_bench_f64_512.json'),('fp32 IISO + DIF 2',f'{R}_bench_f32_iiso_dif2_512.json'),('ref f32',f'{R}_bench_reference_f32_512.json'),('ref f64',f'{R}_bench_reference_f64_512.json')]:
    d=L(f); r=d.get('roofline') or {}
    print(f"| {name} | {d['value']:,.0f} | {d['ms_per_step']:.4f} | {d['e2e']['value']:,.0f} ({d['steps']} steps) | {r.get('kernel_ms_per_step',0)*1e3:.1f} | {r.get('achieved',0):.0f} | {r.get('frac',0):.2f} | {r.get('achieved',0)/8000:.2f} |")
d=L(f'{R}_bench_f32_dif2_512.json')
for v in d['variants']: print(f"| {v['variant']} | {v['value']:,.0f} | {v['ms_per_step']:.4f} | {v['kernel']} | {v['roofline_achieved_gbs']:.0f} | {v['roofline_frac']:.2f} |")
print(d['cpu_baseline']); print(d['clocks'])
for e in ents: print(e['dtype'],e['update_type'],e['dif_order'],round(e['dram_bytes_per_launch']/1e9,3))
PY
