#!/bin/bash
# Runs on the GPU box (under gpurun): the round's evidence into gpurun_out/final/.
#   bench lines (BASELINE configs), ncu launch list of the bench command, one `ncu --set full`
#   capture per hot kernel, compute-sanitizer runs.  tools/ncu_summary.py / ncu_sass_hot.py turn
#   the .ncu-rep files into the text committed under profiles/.
set -u
O=gpurun_out/final
mkdir -p $O
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_configs2_65536ch.json 2> $O/bench_configs2.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference_arm.json 2> $O/bench_reference.err
python bench.py --channels 16384 --steps 20 --warmup 5 > $O/bench_16384ch.json 2> $O/bench_16384.err
python bench.py --channels 4096 --workload corr_msk --steps 20 --warmup 5 > $O/bench_configs1_4096ch_corr_msk.json 2> $O/bench_configs1.err
python bench.py --template reference --channels 16384 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_16384ch_reference_template_1120taps.json 2> $O/bench_l1120.err
python tools/bench_rx.py --sources 2048 --replay > $O/rx_2048src.json 2> $O/rx_2048.err
python tools/bench_rx.py --sources 8192 --replay --no-e2e > $O/rx_8192src.json 2> $O/rx_8192.err
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -s 16 -c 48 --csv --log-file $O/launches_65536ch.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-coherent > $O/ncu_launches.log 2>&1
for k in k_corr_fft k_mix_agc512 k_sqfft; do
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -o $O/ncu_$k python tools/prof_corr.py 4096 > $O/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k "regex:^k_msk$" -s 2 -c 1 -o $O/ncu_k_msk_65536ch python tools/prof_corr.py 65536 > $O/ncu_k_msk_65536.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^k_msk$" -s 2 -c 1 -o $O/ncu_k_msk_16384ch python tools/prof_corr.py 16384 > $O/ncu_k_msk_16384.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck_smoke.log 2>&1
# the full-occupancy paths on a small batch: 48-sample ring, paired fetch, bits written by the loop
B200AIS_MSK_KIND=2 B200AIS_FUSE_TAIL_MIN_CH=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck_smoke_kind2_fused.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_racecheck_smoke.log 2>&1
ls -la $O
