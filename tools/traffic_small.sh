# DRAM-traffic capture (k_extend launches of one 1-spp pass) for the two small workloads, then their bench lines.
# RAYLIB_B200_FUSED=1 keeps the kernel-per-stage path for the capture (a 1-spp pass of these frames would otherwise be one
# cooperative launch with no separate k_extend to name).
TAG=${TAG:-r02fin}; OUT=gpurun_out; NCU="ncu --clock-control none"
for w in random_spheres_640x360_16spp_d5:5 cornell_1920x1080_64spp_d8:8; do
	name=${w%%:*}; depth=${w##*:}
	RAYLIB_B200_FUSED=1 timeout ${CAPTURE_TIMEOUT:-120} $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum -k regex:k_extend -c $depth --csv \
		--log-file $OUT/${TAG}_traffic_${name}.csv python bench.py --workload $name --spp 1 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_traffic_${name}.json 2> $OUT/${TAG}_traffic_${name}.err
	python tools/ncu_traffic.py $OUT/${TAG}_traffic_${name}.csv $OUT/${TAG}_traffic_${name}.json >> $OUT/${TAG}_traffic.log 2>&1
done
cp profiles/traffic.json $OUT/traffic.json
[ -n "$SKIP_BENCH" ] && { cat $OUT/${TAG}_traffic.log; exit 0; }
for w in random_spheres_640x360_16spp_d5:spheres:40 cornell_1920x1080_64spp_d8:cornell:6; do
	name=${w%%:*}; rest=${w#*:}
	timeout 150 python bench.py --workload $name --steps ${rest##*:} --warmup 5 > $OUT/${TAG}_bench_${rest%%:*}.json 2> $OUT/${TAG}_bench_${rest%%:*}.err
done
cat $OUT/${TAG}_traffic.log
