# One GPU: tests, bench (both arms), ncu launch list of the bench command. Usage: bash tools/round_bench.sh <tag>
TAG=${1:-rXX}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --dtype bf16 --no-cpu-baseline > gpurun_out/${TAG}_bench_bf16.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launches.log 2>&1
cat gpurun_out/${TAG}_bench.json
