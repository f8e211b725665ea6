#!/bin/bash
# here (no GPU): turn the files tools/r2_evidence.sh left in gpurun_out/ into the tracked summaries under profiles/
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles; D=3d-reconstruction-detection_b200/csrc/depth.o
cp $G/r2_bench_default.json $G/r2_bench_ground.json $G/r2_bench_reference.json $G/r2_launches.csv $G/r2_rows_mixture.txt $P/
python tools/launch_summary.py $G/r2_launches.csv > $P/r2_launches_summary.txt
python profiles/ncu_summary.py $G/r2_all.ncu-rep > $P/r2_kernels_full.txt
python profiles/ncu_summary.py $G/r2_unproject.ncu-rep > $P/r2_unproject_full.txt
idx() { ncu -i $G/r2_all.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; i=h.index('Kernel Name')
for n,row in enumerate(r[2:]):
    if '$1' in row[i]: print(n); break"; }
python tools/ncu_lines.py $G/r2_all.ncu-rep $D hv_pass_kernelINS_11DepthSourceELi0 6.0 $(idx "DepthSource, 0>") > $P/r2_insert_lines.txt 2>&1
python tools/ncu_lines.py $G/r2_all.ncu-rep $D hv_pass_kernelINS_11DepthSourceELi1 6.0 $(idx "DepthSource, 1>") > $P/r2_lookup_lines.txt 2>&1
python tools/ncu_lines.py $G/r2_all.ncu-rep $D hv_emit_kernelINS_11DepthSource 5.0 $(idx hv_emit) > $P/r2_emit_lines.txt 2>&1
python tools/ncu_lines.py $G/r2_all.ncu-rep $D hv_post_kernelINS_11DepthSource 3.0 $(idx hv_post) > $P/r2_post_lines.txt 2>&1
python - <<'PY'
import json,sys,subprocess,csv,io,importlib.util
spec=importlib.util.spec_from_file_location('bench','bench.py'); b=importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
out=subprocess.run(["ncu","-i","gpurun_out/r2_all.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out))); h=rows[0]; u=rows[1]
i_n=h.index("Kernel Name"); i_r=h.index("dram__bytes_read.sum"); i_w=h.index("dram__bytes_write.sum")
tob=lambda v,unit: float(v.replace(",",""))*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[unit]
ins=[]; kern={}
for r in rows[2:]:
    tot=tob(r[i_r],u[i_r])+tob(r[i_w],u[i_w])
    if "hv_pass_kernel" in r[i_n] and "1>" in r[i_n]:
        kern["hv_pass_kernel<DepthSource,1>"]={"dram_bytes_per_launch":tot,"launches":1}
    elif "hv_pass_kernel" in r[i_n] and "0>" in r[i_n]:
        ins.append(tot)
    elif "hv_emit_kernel" in r[i_n]:
        kern["hv_emit_kernel"]={"dram_bytes_per_launch":tot,"launches":1}
if ins:   # the insert rounds of one step (the capture holds one step): average per launch, like bench.py's per-launch time
    kern["hv_pass_kernel<DepthSource,0>"]={"dram_bytes_per_launch":sum(ins)/len(ins),"launches":len(ins),"note":"average over the step's insert rounds; round 1 alone: %.0f bytes" % ins[0]}
d={"kernels":kern,
   "capture":"profiles/r2_kernels_full.txt: ncu --set full --clock-control none, the launches of one step over 64 frames (bench.py --profile-only, RD3_STREAMS=1)",
   "source_hash":b._source_hash()}
json.dump(d,open('profiles/dominant_kernel_traffic.json','w'),indent=1); print(d)
PY
