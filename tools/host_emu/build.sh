#!/bin/sh
# Developer-only: single-threaded host emulation of one CTA of the CUDA solver, for debugging the
# algorithm on machines without a GPU.  NOT built by __graft_entry__.build(), NOT loaded by the package.
set -e
cd "$(dirname "$0")/../.."
g++ -O2 -g -std=c++17 -fPIC -shared -DOBCA_HOST_EMU -Iinclude -Iconflict_rez_b200/csrc \
    -x c++ conflict_rez_b200/csrc/obca_api.cu -o tools/host_emu/libobca_hostemu.so
echo built tools/host_emu/libobca_hostemu.so
