#!/bin/bash
# tools/gpu_record.sh TAG -- the round's measurement of record on one B200 (run through gpurun): GPU tests, the default bench
# line, the reference arm, then -- each only after its un-profiled command exited 0 -- the ncu launch list of the same bench
# command and one `--set full` capture of the two dominant kernels.  Outputs under gpurun_out/TAG_*.
tag=${1:-rec}
out=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $out/${tag}_gpu_tests.log; echo "tests exit ${PIPESTATUS[0]}" > $out/${tag}_status.txt
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?" >> $out/${tag}_status.txt
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; echo "reference arm exit $?" >> $out/${tag}_status.txt
python bench.py --only --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --only --no-cpu-baseline --steps 2 --warmup 1 > $out/${tag}_ncu_launch.log 2>&1
echo "ncu launch list exit $?" >> $out/${tag}_status.txt
ncu --set full --clock-control none --import-source on -k regex:k2_band -s 2 -c 1 -o $out/${tag}_k2_band -f \
    python bench.py --only --no-cpu-baseline --steps 1 --warmup 1 > $out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_leaf -s 2 -c 1 -o $out/${tag}_k1_leaf -f \
    python bench.py --only --no-cpu-baseline --steps 1 --warmup 1 >> $out/${tag}_ncu_full.log 2>&1
echo "ncu full exit $?" >> $out/${tag}_status.txt
cat $out/${tag}_status.txt $out/${tag}_gpu_tests.log
tail -c 600 $out/${tag}_bench.json
