# One GPU call after a change to the tree pipeline: parity suite first (stop when red), then the DRAM-traffic capture of
# the headline workload, the default bench line, and -- while the call's time lasts -- the traffic of the two other large workloads.
TAG=${TAG:-r02q8}; OUT=gpurun_out; NCU="ncu --clock-control none"; T0=$(date +%s); LIMIT=${LIMIT:-250}
left() { echo $(( LIMIT - ($(date +%s) - T0) )); }
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; tail -2 $OUT/${TAG}_tests.log
grep -q " passed" $OUT/${TAG}_tests.log && ! grep -q "failed\|error" $OUT/${TAG}_tests.log || { echo "PARITY RED"; exit 1; }
rm -f $OUT/${TAG}_traffic.log
capture() {
	name=$1; depth=$2
	timeout $(left) $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum -k regex:k_extend -c $depth --csv \
		--log-file $OUT/${TAG}_traffic_${name}.csv python bench.py --workload $name --spp 1 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_traffic_${name}.json 2> $OUT/${TAG}_traffic_${name}.err \
	&& python tools/ncu_traffic.py $OUT/${TAG}_traffic_${name}.csv $OUT/${TAG}_traffic_${name}.json >> $OUT/${TAG}_traffic.log 2>&1
	cp profiles/traffic.json $OUT/traffic.json
}
capture scatter10M_3840x2160_256spp_d8 8
timeout $(left) python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
python -c "
import json; d=json.loads(open('$OUT/${TAG}_bench_default.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['e2e']['value'], d['roofline']['dram_frac'], d['roofline']['device_tests_per_ray'])"
[ $(left) -gt 45 ] && capture textured2M_1920x1080_64spp_d8 8
[ $(left) -gt 30 ] && capture grid1M_1920x1080_16spp_d2 2
cat $OUT/${TAG}_traffic.log; echo "time left $(left)"
