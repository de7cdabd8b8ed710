#!/bin/bash
# Round-end measurement batch on the GPU box:  bash tools/profile_round.sh <tag>
# Order matters: the DRAM-traffic capture comes first (tools/ncu_traffic.py writes profiles/traffic.json with the hash of the
# device sources), so that the bench lines taken afterwards carry roofline.dram_frac measured on THIS build.  Bench lines
# are never taken under a profiler.
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "radiance|depth|bit-identical|passed|failed|rror|config|primary" | tail -80 > $OUT/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
# DRAM / L2 bytes of the k_extend launches of one 1-spp pass (8 launches at depth 8) -> profiles/traffic.json
for w in scatter10M_3840x2160_256spp_d8:8 textured2M_1920x1080_64spp_d8:8 grid1M_1920x1080_16spp_d2:2; do
	name=${w%%:*}; depth=${w##*:}
	$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum -k regex:k_extend -c $depth --csv \
		--log-file $OUT/${TAG}_traffic_${name}.csv python bench.py --workload $name --spp 1 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_traffic_${name}.json 2> $OUT/${TAG}_traffic_${name}.err
	python tools/ncu_traffic.py $OUT/${TAG}_traffic_${name}.csv $OUT/${TAG}_traffic_${name}.json >> $OUT/${TAG}_traffic.log 2>&1
done
cp profiles/traffic.json $OUT/traffic.json
# bench lines (no profiler attached): the headline, the reference arm, the other four BASELINE configurations -- all WITH cpu_baseline
python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err || exit 1
python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
# (millisecond frames get more steps: two frames of 2 ms say more about the clock ramp than about the renderer)
for w in random_spheres_640x360_16spp_d5:spheres:40 cornell_1920x1080_64spp_d8:cornell:6 grid1M_1920x1080_16spp_d2:grid:20 textured2M_1920x1080_64spp_d8:textured:3; do
	name=${w%%:*}; rest=${w#*:}
	python bench.py --workload $name --steps ${rest##*:} --warmup 5 > $OUT/${TAG}_bench_${rest%%:*}.json 2> $OUT/${TAG}_bench_${rest%%:*}.err
done
# launch list of the default command, then --set full captures (second pass, second bounce: incoherent, binned rays)
$NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file $OUT/${TAG}_launches_bench_default.csv \
	python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
$NCU --set full --import-source on -k regex:k_extend -s 9 -c 1 -f -o $OUT/${TAG}_extend_c4 \
	python bench.py --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > $OUT/${TAG}_ncu_extend.log 2>&1
# the same kernel on camera rays (first launch: coherent, L1-resident nodes) -- what the traversal does when memory is not in the way
$NCU --set full --import-source on -k regex:k_extend -s 0 -c 1 -f -o $OUT/${TAG}_extend_primary_c4 \
	python bench.py --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > $OUT/${TAG}_ncu_extend_primary.log 2>&1
$NCU --set full --import-source on -k regex:k_shadow -s 9 -c 1 -f -o $OUT/${TAG}_shadow_c4 \
	python bench.py --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > $OUT/${TAG}_ncu_shadow.log 2>&1
$NCU --set full --import-source on -k "regex:k_bin_scatter|k_bin_scan|k_shade|k_miss|k_raygen|k_accumulate" -s 10 -c 9 -f -o $OUT/${TAG}_stages_c4 \
	python bench.py --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > $OUT/${TAG}_ncu_stages.log 2>&1
for r in extend extend_primary shadow stages; do
	python tools/ncu_summary.py $OUT/${TAG}_${r}_c4.ncu-rep > $OUT/${TAG}_${r}_scatter10M.txt 2>> $OUT/${TAG}_summary.err
done
ls -la $OUT | grep ${TAG}_ | awk '{print $5, $9}'
