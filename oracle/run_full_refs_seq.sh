#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  A second full-size anchor: init flow + reference global_faldoi runs (methods 0 and 4)
# for another example sequence, frames copied next to the outputs so that the directory travels.
#   oracle/run_full_refs_seq.sh final/hard
set -x
SEQ=${1:-final/hard}
cd "$(dirname "$0")/.."
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib12
R=oracle/_ref
D=$R/data/$(echo "$SEQ" | tr / _)
E=/root/reference/example_data/$SEQ
oracle/make_init_flow.sh "$SEQ" 0 > /dev/null 2>&1
cp "$E"/frame_000[1-4].png "$D"/
cp "$E"/gt/frame_0002.flo "$D"/gt_frame_0002.flo 2>/dev/null
( time $R/global_faldoi $D/ims.txt $D/rg.flo $D/var_m0.flo -m 0 -w 5 -verbose 1 ) > $D/log_m0.txt 2>&1
( time $R/global_faldoi $D/ims.txt $D/rg.flo $D/var_m4.flo -m 4 -w 5 -verbose 1 ) > $D/log_m4.txt 2>&1
# TV-CSAD without the data race of tvcsad_getP's error sum: one thread, one warp (about 10 minutes)
( time OMP_NUM_THREADS=1 $R/global_faldoi $D/ims.txt $D/rg.flo $D/var_m4_w1_t1.flo -m 4 -w 1 -verbose 1 ) > $D/log_m4_w1_t1.txt 2>&1
echo ALLDONE
