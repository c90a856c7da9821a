#!/bin/bash
# One GPU-box session: every step under its own timeout, every log under gpurun_out/ (merged back by gpurun).
# Usage (on the GPU box): bash tools/gpu_session.sh <tag> [steps...]   steps: test bench probe diag cfg5 launches ncu
TAG=${1:-s}; shift
STEPS=${@:-test bench probe diag cfg5 launches}
O=gpurun_out; mkdir -p $O
for S in $STEPS; do
  T0=$(date +%s)
  case $S in
    test)     timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -s -p no:cacheprovider > $O/${TAG}_test.log 2>&1 ;;
    testfast) timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -s -p no:cacheprovider -k "not test_layer and not 25920 and not fused_frame_vs_golden and not colorvidnet_module and not warpnet_module" > $O/${TAG}_testfast.log 2>&1 ;;
    bench)    timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err ;;
    probe)    timeout 120 tools/probes/umma_rowshift_probe > $O/${TAG}_probe.log 2>&1 ;;
    diag)     timeout 300 python tools/diag_corr_candidates.py > $O/${TAG}_diag.log 2>&1 ;;
    cfg5)     timeout 600 python tools/config5_corr_microbench.py --out $O/config5_r2.jsonl > $O/${TAG}_cfg5.log 2>&1 ;;
    cfg4)     timeout 600 bash tools/run_config4.sh > $O/${TAG}_cfg4.log 2>&1 ;;
    extra)    timeout 600 python tools/extra_configs.py > $O/${TAG}_extra.log 2>&1 ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${TAG}_launches.csv \
                python bench.py --steps 4 --warmup 3 --cpu-sample 0 --sustain-s 0 --clip-frames 0 > $O/${TAG}_launches.out 2>&1 ;;
    ncu)      timeout 1200 bash tools/ncu_captures.sh $TAG > $O/${TAG}_ncu.log 2>&1 ;;
    rs)       DVC_TEST_ROWSHARE=1,2 timeout 600 python -m pytest tests/test_gpu_rowshare.py -m gpu -q --maxfail=400 -s -p no:cacheprovider > $O/${TAG}_rs.log 2>&1 ;;
    layers)   timeout 300 python tools/conv_layer_bench.py --out $O/conv_layers_r2.jsonl > $O/${TAG}_layers.log 2>&1 ;;
    benchrs)  timeout 600 python bench.py --steps 20 --warmup 5 --tc-rowshare 1 --cpu-sample 0 --clip-frames 0 > $O/${TAG}_benchrs.json 2> $O/${TAG}_benchrs.err ;;
    bench3)   timeout 600 python bench.py --steps 20 --warmup 5 --clip-astreams 2 --cpu-sample 0 --clip-frames 0 > $O/${TAG}_bench3.json 2> $O/${TAG}_bench3.err ;;
    bench3rs) timeout 600 python bench.py --steps 20 --warmup 5 --clip-astreams 2 --tc-rowshare 1 --cpu-sample 0 --clip-frames 0 > $O/${TAG}_bench3rs.json 2> $O/${TAG}_bench3rs.err ;;
    benchN)   NG=$(nvidia-smi -L | wc -l); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29711 \
                bench.py --gpus $NG --steps 20 --warmup 5 --cpu-sample 0 > $O/${TAG}_bench${NG}gpu.json 2> $O/${TAG}_bench${NG}gpu.err ;;
    sanitize) for tool in memcheck racecheck synccheck initcheck; do
                timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize_small.py > $O/${TAG}_sanitize_$tool.log 2>&1
                echo "  sanitizer $tool rc=$? $(grep -c 'ERROR SUMMARY' $O/${TAG}_sanitize_$tool.log)" >> $O/${TAG}_steps.log; tail -n 3 $O/${TAG}_sanitize_$tool.log >> $O/${TAG}_steps.log
              done ;;
    layersx)  DVC_LAYER_EXPERIMENTS=1 timeout 300 python tools/conv_layer_bench.py > $O/${TAG}_layersx.log 2>&1 ;;
    ab_alt)   L=deep-exemplar-based-video-colorization_b200/lib; cp $L/libdvc.so /tmp/libdvc_main.so; cp $L/libdvc_alt.so $L/libdvc.so;
              timeout 300 python tools/conv_layer_bench.py > $O/${TAG}_layers_alt.log 2>&1;
              timeout 300 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --clip-frames 0 --sustain-s 0 > $O/${TAG}_bench_alt.json 2> $O/${TAG}_bench_alt.err;
              cp /tmp/libdvc_main.so $L/libdvc.so ;;
    convtest) timeout 240 python -m pytest tests/test_gpu_conv_layers.py -m gpu -q -x -p no:cacheprovider > $O/${TAG}_convtest.log 2>&1 ;;
    testq)    timeout 420 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/${TAG}_test.log 2>&1 ;;
    folder)   timeout 300 python - > $O/${TAG}_folder.log 2>&1 <<'PYEOF'
import os, subprocess, sys
import numpy as np
from PIL import Image
rng = np.random.default_rng(0)
os.makedirs("/tmp/clip", exist_ok=True)
base = np.kron(rng.integers(0, 255, (54, 96, 3)), np.ones((10, 10, 1))).astype(np.uint8)      # 540 x 960 blocky content
for t in range(4):
    img = np.clip(np.roll(base, 7 * t, axis=1).astype(np.int32) + rng.integers(-12, 13, base.shape), 0, 255).astype(np.uint8)
    Image.fromarray(img).convert("L").convert("RGB").save(f"/tmp/clip/frame{t:03d}.png")       # grayscale frames
Image.fromarray(base).save("/tmp/ref.png")
rc = subprocess.call([sys.executable, "tools/colorize_folder.py", "--clip", "/tmp/clip", "--ref", "/tmp/ref.png", "--out", "/tmp/out",
                      "--seeded-weights"])
outs = sorted(os.listdir("/tmp/out")) if os.path.isdir("/tmp/out") else []
print("rc", rc, "outputs", outs)
for o in outs:
    a = np.asarray(Image.open(os.path.join("/tmp/out", o)))
    print(o, a.shape, a.dtype, float(a.mean()), float(a.std()))
assert rc == 0 and len(outs) == 4
PYEOF
              ;;
    smoke)    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1 ;;
    rsq)      DVC_TEST_ROWSHARE=1 timeout 150 python -m pytest tests/test_gpu_rowshare.py -m gpu -q -x -p no:cacheprovider > $O/${TAG}_rsq.log 2>&1 ;;
    multi)    timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider > $O/${TAG}_multi.log 2>&1 ;;
    *) echo "unknown step $S" ;;
  esac
  RC=$?
  echo "step $S rc=$RC $(( $(date +%s) - T0 ))s" >> $O/${TAG}_steps.log
  if [ "$S" = "convtest" ] && [ $RC -ne 0 ]; then echo "convtest failed: stopping the session" >> $O/${TAG}_steps.log; break; fi
done
tail -n 12 $O/${TAG}_steps.log
for f in $O/${TAG}_test.log $O/${TAG}_testfast.log; do [ -f $f ] && tail -n 6 $f; done
exit 0
