#!/bin/bash
# First GPU call of round 2: runs everything round 1 wrote after its GPU budget was spent (the widened rows, the
# mini-training PSNR comparison, PeerAdam) and the A/B of the fused L1 loss.  One GPU unless NGPU is set.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/r2_first_call.sh'
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'NGPU=2 bash profiles/r2_first_call.sh'
set -x
mkdir -p gpurun_out
O=gpurun_out
if [ "${NGPU:-1}" -gt 1 ]; then
  FGS_TEST_PEER_ADAM=1 timeout 600 python -m pytest tests/test_peer_adam.py -m gpu -x -q -s > $O/r2_peer_adam_test.txt 2>&1
  for flag in "" "--flat-allreduce" "--peer-adam"; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $NGPU --steps 100 --warmup 10 $flag > "$O/r2_bench_${NGPU}gpu${flag}.json" 2> "$O/r2_bench_${NGPU}gpu${flag}.err"
  done
  exit 0
fi
timeout 900 python -m pytest tests/test_widen_loss.py tests/test_widen_densify.py tests/test_widen_training.py -m gpu -q -s > $O/r2_widen_tests.txt 2>&1
tail -5 $O/r2_widen_tests.txt
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/r2_bench_default.json 2> $O/r2_bench_default.err
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --fused-loss > $O/r2_bench_fusedloss.json 2> $O/r2_bench_fusedloss.err
timeout 600 python profiles/bench_widen.py > $O/r2_widen_bench.json 2> $O/r2_widen_bench.err
# per-launch list of the widened kernels (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ssim|l1_kernel|gather_rows|plan_|adc_|morton|bbox' \
  -c 60 --csv --log-file $O/r2_widen_launches.csv python profiles/bench_widen.py > /dev/null 2>&1
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --overlap-adam > $O/r2_bench_overlap.json 2> $O/r2_bench_overlap.err
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --overlap-adam --fused-loss > $O/r2_bench_overlap_fusedloss.json 2> $O/r2_bench_overlap_fusedloss.err
timeout 600 python -m pytest tests/test_widen_overlap_gpu.py -m gpu -q -s > $O/r2_overlap_test.txt 2>&1
timeout 900 python -m pytest tests/test_widen_fourd.py -m gpu -q -s > $O/r2_fourd_test.txt 2>&1
timeout 600 python profiles/bench_fourd.py > $O/r2_fourd_bench.json 2> $O/r2_fourd_bench.err
# one --set full capture of the two heaviest new kernels (read here with ncu -i ... --page raw --csv)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ssim_stats|ssim_grad|gather_rows' -c 6 -o $O/r2_widen_full \
  python profiles/bench_widen.py > /dev/null 2>&1
