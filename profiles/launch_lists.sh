mkdir -p gpurun_out/r2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum
run() {  # tag, bench args
  tag=$1; shift
  timeout 400 python bench.py "$@" --kernel-only --steps 10 --warmup 3 2>/dev/null | grep "^{" > gpurun_out/r2/kernel_only_v31_$tag.json
  timeout 500 ncu --metrics $M --clock-control none -c 120 --csv --log-file gpurun_out/r2/launches_$tag.csv python bench.py "$@" --kernel-only --steps 4 --warmup 3 > /dev/null 2>&1
  python -c "
import json,sys
d=json.load(open('gpurun_out/r2/kernel_only_v31_$tag.json')); print('$tag', d['kernel_ms'], d['scan_ms'], d['GBps'], d['stats']['n_hits'], d['stats']['n_records'])"
}
run C2 --workload C2
run C4 --workload C4
run C3 --workload C3
run C5 --workload C5 --scale 0.02
run C5full --workload C5 --catalogue full --scale 0.01
