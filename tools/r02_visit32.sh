# round 2, GPU visit 32: few rays per cell (the reference's decks use 3 ... 100) -- sorted vs unsorted blocked vs linear kernel
T=r02v32
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-sweeps --steps 5"
for rays in 3 5 10 25 50; do for bm in 3 2 0; do
$B --rays $rays --blocked-march $bm > $O/${T}_r${rays}_bm${bm}.json 2>> $O/${T}_err.log; python -c "import json;d=json.load(open('$O/${T}_r${rays}_bm${bm}.json'));print('rays',$rays,'bm',$bm,'%.3e'%d['value'],'%.3f ms'%d['trace_ms_per_step'])"
done; done
tail -3 $O/${T}_err.log
