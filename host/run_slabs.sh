#!/bin/bash
# Launcher of the C++ z-slab driver: one process per GPU on this node, no MPI.
#   host/run_slabs.sh N [slab_driver options]      e.g.  host/run_slabs.sh 8 --nx 2048 --ny 2048 --nzl 128 --steps 10
# The ranks meet through a file in a fresh directory (rank 0 writes the NCCL id there).
N=${1:-1}; shift
HERE="$(cd "$(dirname "$0")" && pwd)"
DIR=$(mktemp -d)
export SW4B200_ID_FILE=$DIR/nccl_id WORLD_SIZE=$N
pids=""
for r in $(seq 0 $((N-1))); do
   RANK=$r LOCAL_RANK=$r "$HERE/_build/slab_driver" "$@" &
   pids="$pids $!"
done
rc=0
for p in $pids; do wait $p || rc=1; done
rm -rf "$DIR"
exit $rc
