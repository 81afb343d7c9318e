#!/usr/bin/env bash
# Round-2 GPU session: tests, bench, launch list.  Usage (on the GPU box, from the repo root):  bash tools/gpu_r2.sh <tag> [stages...]
# stages: tests bench launches layers (default: all four)
set -u
TAG=${1:-r2a}; shift || true
STAGES=${*:-tests bench launches layers}
mkdir -p gpurun_out
for st in $STAGES; do
  case $st in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -5 gpurun_out/${TAG}_tests.log;;
    quick)
      timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_curves_gpu.py > gpurun_out/${TAG}_tests.log 2>&1; echo "quick tests rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -5 gpurun_out/${TAG}_tests.log;;
    curves)
      DCV_CURVE_DUMP=gpurun_out/${TAG}_curves timeout 1500 python -m pytest tests/test_curves_gpu.py -m gpu -q -s > gpurun_out/${TAG}_curves.log 2>&1; echo "curves rc=$?" | tee -a gpurun_out/${TAG}_rc.log; grep -v "^DEBUG\|^INFO" gpurun_out/${TAG}_curves.log | tail -30;;
    imgtest)
      timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_nets_gpu.py tests/test_timed_path_gpu.py -m gpu -q -s -k "img_conv or full_width or generators_match or graph or golden" > gpurun_out/${TAG}_imgtest.log 2>&1; echo "imgtest rc=$?" | tee -a gpurun_out/${TAG}_rc.log; grep -v "^DEBUG\|^INFO" gpurun_out/${TAG}_imgtest.log | tail -25;;
    imglayers)
      timeout 600 python tools/layer_bench.py mug-depth 32 > gpurun_out/${TAG}_layers_img.md 2>&1; echo "imglayers rc=$?" | tee -a gpurun_out/${TAG}_rc.log; head -12 gpurun_out/${TAG}_layers_img.md;;
    ncuimg)
      python tools/probe_img.py > gpurun_out/${TAG}_probe_img.log 2>&1
      echo "probe only"; echo "ncuimg rc=$?" | tee -a gpurun_out/${TAG}_rc.log; cat gpurun_out/${TAG}_probe_img.log;;
    dp)
      # data-parallel A/B on the GPUs of this box: side-stream overlap with reserved SMs (default) against in-stream reduction
      NG=${NG:-2}
      for mode in ${MODES:-0 1 2}; do
        DCV_DP_OVERLAP=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 --no-roofline > gpurun_out/${TAG}_bench_${NG}gpu_overlap${mode}.json 2> gpurun_out/${TAG}_bench_${NG}gpu_overlap${mode}.err; echo "dp $NG overlap=$mode rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -c 400 gpurun_out/${TAG}_bench_${NG}gpu_overlap${mode}.json
      done;;
    dpconfigs)
      NG=${NG:-8}
      for c in surreal-depth1 isogd-flow surreal-segm; do
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 --no-roofline --config $c > gpurun_out/${TAG}_bench_${NG}gpu_${c}.json 2> gpurun_out/${TAG}_bench_${NG}gpu_${c}.err; echo "dp $NG $c rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -c 300 gpurun_out/${TAG}_bench_${NG}gpu_${c}.json
      done;;
    ncutop)
      # one `ncu --set full` capture per dominant launch (B200_PROFILING.md recipe), after the plain run exited 0
      python tools/probe_layers.py vdis_main1_dgrad up5_fwd vdis_main1_fwd vdis_main1_wgrad down0_wgrad > gpurun_out/${TAG}_probe_layers.log 2>&1 && \
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_pers|wgrad_tc" -c 10 -o gpurun_out/${TAG}_top python tools/probe_layers.py vdis_main1_dgrad up5_fwd vdis_main1_fwd vdis_main1_wgrad down0_wgrad > gpurun_out/${TAG}_ncutop.log 2>&1; echo "ncutop rc=$?" | tee -a gpurun_out/${TAG}_rc.log; cat gpurun_out/${TAG}_probe_layers.log;;
    tf32)
      timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_nets_gpu.py -m gpu -q -s -k "tf32 or full_width" > gpurun_out/${TAG}_tf32.log 2>&1; echo "tf32 rc=$?" | tee -a gpurun_out/${TAG}_rc.log; grep -v "^DEBUG\|^INFO" gpurun_out/${TAG}_tf32.log | grep "tf32\|passed\|failed\|Error\|error" | tail -30;;
    sanitize)
      # memcheck + racecheck over the kernels added this round (image-side mma.sync kernels, in-kernel BatchNorm finalize, head + loss,
      # transposing reduction, TMA reduce-add): small shapes, a few tests each
      for tool in memcheck racecheck; do
        timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "img_conv or outconv or batchnorm or losses or head_with or accumulate or tc_conv2d_64 or tc_convT_128_64" > gpurun_out/${TAG}_sanitize_${tool}.log 2>&1; echo "sanitize $tool rc=$?" | tee -a gpurun_out/${TAG}_rc.log; grep -c "ERROR SUMMARY\|Error\|RACECHECK" gpurun_out/${TAG}_sanitize_${tool}.log; tail -4 gpurun_out/${TAG}_sanitize_${tool}.log
      done;;
    bench)
      timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -c 3000 gpurun_out/${TAG}_bench.json;;
    benchfast)
      timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -c 3000 gpurun_out/${TAG}_bench.json;;
    benchpdl)
      timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-roofline --tune pdl=1 > gpurun_out/${TAG}_bench_pdl.json 2> gpurun_out/${TAG}_bench_pdl.err; echo "bench pdl rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -c 700 gpurun_out/${TAG}_bench_pdl.json;;
    benchimg)
      DCV_NO_IMG_CONV=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-roofline > gpurun_out/${TAG}_bench_img.json 2> gpurun_out/${TAG}_bench_img.err; echo "bench img rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -c 700 gpurun_out/${TAG}_bench_img.json;;
    refarm)
      timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_ref.err; echo "refarm rc=$?" | tee -a gpurun_out/${TAG}_rc.log; cat gpurun_out/${TAG}_bench_reference_arm.json;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu --no-eager --no-roofline > gpurun_out/${TAG}_ncu.log 2>&1; echo "launches rc=$?" | tee -a gpurun_out/${TAG}_rc.log;;
    layers)
      timeout 600 python tools/layer_bench.py mug-depth 32 > gpurun_out/${TAG}_layers.md 2>&1; echo "layers rc=$?" | tee -a gpurun_out/${TAG}_rc.log; head -30 gpurun_out/${TAG}_layers.md;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -3 gpurun_out/${TAG}_smoke.log;;
    configs)
      for c in surreal-depth1 isogd-flow surreal-segm; do
        timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --no-roofline --config $c > gpurun_out/${TAG}_bench_${c}.json 2> gpurun_out/${TAG}_bench_${c}.err; echo "bench $c rc=$?" | tee -a gpurun_out/${TAG}_rc.log; tail -c 600 gpurun_out/${TAG}_bench_${c}.json
      done;;
  esac
done
