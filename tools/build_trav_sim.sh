#!/bin/bash
# Builds tools/trav_sim (host-only development tool) against the product library and the scene client library.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
g++ -std=c++17 -O2 -pthread -I"$ROOT/include" -I"$ROOT/software-raytracing_b200/csrc/host" "$ROOT/tools/trav_sim.cc" "$ROOT/tools/bvh_reinsert.cc" -L"$ROOT/software-raytracing_b200/lib" -L"$ROOT/scenes/lib" \
    -lscenes_b200 -lraylib_b200 -Wl,-rpath,"$ROOT/software-raytracing_b200/lib" -Wl,-rpath,"$ROOT/scenes/lib" -o "$ROOT/tools/trav_sim"
