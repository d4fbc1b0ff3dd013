O=gpurun_out/r02H; mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1; tail -2 $O/pytest.txt
python bench.py > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python bench.py --impl reference > $O/bench_ref.json 2>> $O/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --pages 262144 --steps 2 --warmup 3 --no-cpu --no-alt --no-workloads > $O/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:^compress_kernel -s 3 -c 1 -o $O/prof_c python bench.py --pages 262144 --steps 1 --warmup 3 --no-e2e --no-cpu --no-alt --no-workloads > $O/ncu_c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:decompress -s 3 -c 1 -o $O/prof_d python bench.py --pages 262144 --steps 1 --warmup 3 --no-e2e --no-cpu --no-alt --no-workloads > $O/ncu_d.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:^compress_kernel -s 3 -c 1 -o $O/prof_c_text python bench.py --pages 65536 --steps 1 --warmup 3 --no-e2e --no-cpu --no-alt --no-workloads --only text > $O/ncu_ct.log 2>&1
for t in memcheck synccheck racecheck; do timeout 500 compute-sanitizer --tool $t python tools/sanitize_run.py > $O/sanitizer_$t.log 2>&1; echo "$t rc=$?"; tail -1 $O/sanitizer_$t.log; done
python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['compress_gbs'], d['decompress_gbs'], d['e2e']['value'], d['e2e_concurrent']['value'], d['e2e_pageable']['value'], d['roofline']['frac'])"
ls $O
timeout 400 python tools/fuzz_gpu.py 300 2>&1 | tail -1 | tee $O/fuzz_soak.txt
