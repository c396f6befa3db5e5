set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r4_pytest.log 2>&1; tail -2 $O/r4_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r4_launches.csv python bench.py --steps 2 --warmup 1 > $O/r4_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_slots_fast --launch-skip 2 -c 1 -f -o $O/r4_slots python bench.py --workload c2 --steps 1 --warmup 1 --no-parity --no-cpu-baseline --seconds 0 --also "" > $O/r4_ncu_slots.log 2>&1
python bench.py > $O/r4_bench_default_c3_n1.json 2> $O/r4_bench_default.err
python bench.py --impl reference > $O/r4_bench_ref_c3.json 2>> $O/r4_bench_default.err
python bench.py --impl reference --workload c2 > $O/r4_bench_ref_c2.json 2>> $O/r4_bench_default.err
python bench.py --workload c4 --also "" > $O/r4_bench_c4_n1.json 2>> $O/r4_bench_default.err
python bench.py --sweep --no-cpu-baseline --seconds 0 --steps 10 > $O/r4_bench_c5_sweep_n1.json 2>> $O/r4_bench_default.err
python bench.py --workload c2v6 --also "" --no-cpu-baseline > $O/r4_bench_c2v6_n1.json 2>> $O/r4_bench_default.err
python bench.py --workload c3h --also "" --no-cpu-baseline > $O/r4_bench_c3h_n1.json 2>> $O/r4_bench_default.err
python bench.py --workload c2 --untiled --also "" --no-cpu-baseline > $O/r4_bench_c2_untiled_n1.json 2>> $O/r4_bench_default.err
python bench.py --workload c2 --also "" --no-cpu-baseline > $O/r4_bench_c2_n1.json 2>> $O/r4_bench_default.err
tail -5 $O/r4_bench_default.err
cut -c1-300 $O/r4_bench_default_c3_n1.json
