#!/bin/bash
# Development (GPU box): run tools/sweep.py with every library variant under software-raytracing_b200/build/variants/
#   bash tools/try_variants.sh "<cfg> <spp> [ENV=...]" ["<cfg> <spp>" ...]
LIB=software-raytracing_b200/lib/libraylib_b200.so
cp $LIB /tmp/base.so
for v in base $(ls software-raytracing_b200/build/variants/*.so 2>/dev/null); do
	name=$(basename $v .so)
	if [ "$v" = base ]; then cp /tmp/base.so $LIB; else cp $v $LIB; fi
	for args in "$@"; do
		python tools/sweep.py $args 2>/dev/null | sed "s/^{/{\"variant\": \"$name\", /"
	done
done
cp /tmp/base.so $LIB
