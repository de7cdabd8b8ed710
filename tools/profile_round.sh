#!/bin/bash
# Round-end measurement batch on the GPU box:  bash tools/profile_round.sh <tag>
# Bench lines first (no profiler attached), then the ncu launch list, the --set full captures and the DRAM-traffic pass.
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err || exit 1
python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
for w in random_spheres_640x360_16spp_d5:spheres cornell_1920x1080_64spp_d8:cornell grid1M_1920x1080_16spp_d2:grid textured2M_1920x1080_64spp_d8:textured; do
	python bench.py --workload ${w%%:*} --no-cpu-baseline > $OUT/${TAG}_bench_${w##*:}.json 2>/dev/null
done
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file $OUT/${TAG}_launches_bench_default.csv \
	python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
# second pass, second bounce (incoherent, binned rays): launch index 9 of k_extend / k_shadow
$NCU --set full --import-source on -k regex:k_extend -s 9 -c 1 -f -o $OUT/${TAG}_extend_c4 \
	python bench.py --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > $OUT/${TAG}_ncu_extend.log 2>&1
$NCU --set full --import-source on -k regex:k_shadow -s 9 -c 1 -f -o $OUT/${TAG}_shadow_c4 \
	python bench.py --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > $OUT/${TAG}_ncu_shadow.log 2>&1
$NCU --set full --import-source on -k "regex:k_bin_scatter|k_bin_scan|k_shade|k_miss|k_raygen|k_accumulate" -s 12 -c 9 -f -o $OUT/${TAG}_stages_c4 \
	python bench.py --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > $OUT/${TAG}_ncu_stages.log 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum -k regex:k_extend -c 8 --csv \
	--log-file $OUT/${TAG}_traffic_c4.csv python bench.py --spp 1 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_traffic_c4.json 2> $OUT/${TAG}_traffic.err
ls -la $OUT | grep ${TAG}_ | awk '{print $5, $9}'
