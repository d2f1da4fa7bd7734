#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  TVL2 anchors for the remaining example sequences: init flow (local_faldoi) +
# reference global_faldoi -m 0, frames copied next to the outputs, bulky intermediates removed.
cd "$(dirname "$0")/.."
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib12
R=oracle/_ref
for SEQ in clean/medium clean/hard final/easy final/medium; do
  D=$R/data/$(echo "$SEQ" | tr / _)
  E=/root/reference/example_data/$SEQ
  oracle/make_init_flow.sh "$SEQ" 0 > /dev/null 2>&1
  cp "$E"/frame_000[1-4].png "$D"/
  cp "$E"/gt/frame_0002.flo "$D"/gt_frame_0002.flo 2>/dev/null
  ( time $R/global_faldoi $D/ims.txt $D/rg.flo $D/var_m0.flo -m 0 -w 5 -verbose 1 ) > $D/log_m0.txt 2>&1
  rm -f $D/sim.tiff $D/s1.flo $D/s2.flo $D/m1.txt $D/m2.txt $D/m1_cut.txt $D/m2_cut.txt
  echo "done $SEQ"
done
echo ALLDONE
