#!/bin/bash
# scope B of SURVEY 8d: wall clock of the drop-in command line (FASTQ files in, gz text files out)
mkdir -p gpurun_out
WD=$(mktemp -d /tmp/pedk_dropin_XXXX)
python - "$WD" <<PY
import os, sys
sys.path.insert(0, '.')
from ped_b200 import api
wd = sys.argv[1]
G, N = 5_000_000, 1_000_000
g = api.synth_genome(1012, G)
t = api.synth_mutate(g, 2012, 1000, 100)
for s, gg, seed in (("C", g, 3012), ("T", t, 4012)):
    os.makedirs(f"{wd}/{s}/read")
    open(f"{wd}/{s}/read/{s}.fastq", "wb").write(api.synth_fastq(gg, seed, N, read_len=150))
print("inputs written", wd)
PY
( time perl perl/ped_kmer.pl target=T,control=C,wd=$WD > gpurun_out/dropin_wall.log 2>&1 ) 2> gpurun_out/dropin_time.txt
echo "rc=$?"; tail -4 gpurun_out/dropin_wall.log; cat gpurun_out/dropin_time.txt | tail -3
wc -l $WD/T/T.kmer.snp; du -sh $WD/C/count $WD/C/lbc $WD/C/sort_uniq | tr '\n' ' '
rm -rf $WD
