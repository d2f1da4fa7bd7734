#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Produces the local_faldoi init flow the north star
# names ("local_faldoi flow as init") for one example sequence, with the
# reference's own binaries (oracle/_ref, ext_bin matchers), following the
# pipeline of scripts_python/faldoi_sift.py.
#
#   oracle/make_init_flow.sh clean/easy [method]   ->  oracle/_ref/data/clean_easy/{ims.txt,rg.flo[,rg_occ.png]}
#
# Only runs where /root/reference exists (~2 min per sequence on 8 cores).
set -euo pipefail
SEQ=${1:-clean/easy}
METHOD=${2:-0}
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${REF:-/root/reference}
R=$HERE/_ref
D=$R/data/$(echo "$SEQ" | tr / _)
mkdir -p "$D"
export LD_LIBRARY_PATH=$R/lib12
E=$REF/example_data/$SEQ
# frames: I0 = frame_0002, I1 = frame_0003, I-1 = frame_0001, I2 = frame_0004 (example_data/README.txt)
printf '%s\n%s\n%s\n%s\n' "$E/frame_0002.png" "$E/frame_0003.png" "$E/frame_0001.png" "$E/frame_0004.png" > "$D/ims.txt"
read -r W H < <(python3 -c "from PIL import Image; im=Image.open('$E/frame_0002.png'); print(im.size[0], im.size[1])")
if [ ! -f "$D/s1.flo" ]; then
  "$REF/ext_bin/sift_cli" "$E/frame_0002.png" -ss_nspo 15 > "$D/d1.txt"
  "$REF/ext_bin/sift_cli" "$E/frame_0003.png" -ss_nspo 15 > "$D/d2.txt"
  "$REF/ext_bin/match_cli" "$D/d1.txt" "$D/d2.txt" > "$D/m1.txt"
  "$REF/ext_bin/match_cli" "$D/d2.txt" "$D/d1.txt" > "$D/m2.txt"
  awk '{print $2,$1,$6,$5}' "$D/m1.txt" > "$D/m1_cut.txt"
  awk '{print $2,$1,$6,$5}' "$D/m2.txt" > "$D/m2_cut.txt"
  "$R/sparse_flow" "$D/m1_cut.txt" "$W" "$H" "$D/s1.flo"
  "$R/sparse_flow" "$D/m2_cut.txt" "$W" "$H" "$D/s2.flo"
fi
if [ "$METHOD" = 8 ]; then
  "$R/local_faldoi" "$D/ims.txt" "$D/s1.flo" "$D/s2.flo" "$D/rg_m8.flo" "$D/sim_m8.tiff" "$D/rg_occ.png" \
      -m 8 -wr 5 -loc_it 3 -max_pch_it 4 -split_img 0 -fb_thresh 2
else
  "$R/local_faldoi" "$D/ims.txt" "$D/s1.flo" "$D/s2.flo" "$D/rg.flo" "$D/sim.tiff" \
      -m "$METHOD" -wr 5 -loc_it 3 -max_pch_it 4 -split_img 0 -fb_thresh 2
fi
rm -f "$D"/d1.txt "$D"/d2.txt
ls -la "$D"
