#!/bin/bash
# full-size reference runs for the named BASELINE configs (clean/easy), ~30 min of CPU
set -x
cd /root/repo
export LD_LIBRARY_PATH=/root/repo/oracle/_ref/lib12
D=oracle/_ref/data/clean_easy
R=oracle/_ref
( time $R/global_faldoi $D/ims.txt $D/rg.flo $D/var_m4.flo -m 4 -w 5 -verbose 1 ) > $D/log_m4.txt 2>&1
( time $R/global_faldoi $D/ims.txt $D/rg.flo $D/var_m7.flo -m 7 -w 5 -verbose 1 ) > $D/log_m7.txt 2>&1
oracle/make_init_flow.sh clean/easy 8 > $D/log_local_m8.txt 2>&1
( time $R/global_faldoi $D/ims.txt $D/rg_m8.flo $D/var_m8.flo $D/rg_occ.png $D/var_m8_occ.png -m 8 -w 5 -glb_iters 400 -verbose 1 ) > $D/log_m8.txt 2>&1
echo ALLDONE
