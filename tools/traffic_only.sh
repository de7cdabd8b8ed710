TAG=r02z; OUT=gpurun_out; NCU="ncu --clock-control none"
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > $OUT/${TAG}_tests_tail.log
rm -f $OUT/${TAG}_traffic.log
for w in scatter10M_3840x2160_256spp_d8:8 textured2M_1920x1080_64spp_d8:8 grid1M_1920x1080_16spp_d2:2; do
	name=${w%%:*}; depth=${w##*:}
	$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum -k regex:k_extend -c $depth --csv \
		--log-file $OUT/${TAG}_traffic_${name}.csv python bench.py --workload $name --spp 1 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_traffic_${name}.json 2> $OUT/${TAG}_traffic_${name}.err
	python tools/ncu_traffic.py $OUT/${TAG}_traffic_${name}.csv $OUT/${TAG}_traffic_${name}.json >> $OUT/${TAG}_traffic.log 2>&1
done
cp profiles/traffic.json $OUT/traffic.json
python bench.py > $OUT/${TAG}_bench_default_recheck.json 2> $OUT/${TAG}_bench_default_recheck.err
cat $OUT/${TAG}_tests_tail.log; python -c "
import json; d=json.loads(open('$OUT/${TAG}_bench_default_recheck.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['dram_frac'], d['roofline']['traffic_source'][:60])"
