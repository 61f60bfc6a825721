#!/bin/bash
# the driver's round-end sequence on the final tree: GPU tests, reference arm, own arm (default flags), smoke
O=gpurun_out/r4s; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; tail -2 $O/gputests.log
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; python -c "
import json;d=json.loads(open('$O/bench_ref.json').readlines()[-1]);print('reference',d['value'],d['cpu_baseline']['cores'])"
python bench.py > $O/bench.json 2> $O/bench.err; tail -c 200 $O/bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python -c "
import json;d=json.load(open('$O/bench.json'));print(round(d['value'],1),round(d['e2e']['value'],1),d['gpu_launches'],d['parity_in_run']['identical'],d['roofline']['onchip'].get('timed_region',{}).get('frac'))
a=d['aux'];print(a['frames_per_s'],a['nv12_ingest']['frames_per_s'],a['native_host_threads']['bgr']['frames_per_s'],a['native_host_threads']['nv12']['frames_per_s'],a['device_resident']['frames_per_s'])
o=a['other_configs_one_stream'];print({k:(v.get('frames_per_s') if isinstance(v,dict) else None) for k,v in o.items()})
print(a['element_shaped_sync_calls'])"
