#!/bin/bash
# BASELINE.json's other configurations on ONE GPU (bench.py size / dtype options); one JSON line each.
mkdir -p gpurun_out
: > gpurun_out/configs.jsonl
run() { timeout 600 python bench.py --no-extras "$@" >> gpurun_out/configs.jsonl 2>> gpurun_out/configs.err; echo "exit $? : $*"; }
run --stencil hori_diff --ni 512 --nj 512 --dtype f64 --steps 200 --warmup 20      # configs[2]
run --stencil hori_diff --ni 512 --nj 512 --dtype f32 --steps 200 --warmup 20      # configs[2]
run --stencil hori_diff --ni 128 --nj 128 --dtype f64 --steps 200 --warmup 20      # configs[0] size
run --stencil hori_diff --ni 1024 --nj 1024 --dtype f64 --steps 50 --warmup 5       # configs[4] tile size
run --stencil vert_adv --ni 1024 --nj 1024 --dtype f64 --steps 50 --warmup 5
run --stencil vert_adv --ni 512 --nj 512 --dtype f64 --steps 100 --warmup 10
run --stencil vert_adv --ni 256 --nj 256 --dtype f32 --steps 200 --warmup 20
run --stencil hori_diff --ni 4096 --nj 4096 --dtype f64 --steps 10 --warmup 3       # configs[3], one GPU
python3 - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('%-44s %9.0f Mpts/s %9.1f us/step  %5.0f GB/s  frac %.3f' % (d['metric'], d['value'], d['ms_per_step'] * 1e3, r['achieved'], r['frac']))
PY
tail -3 gpurun_out/configs.err
