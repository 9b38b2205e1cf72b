package PedK;
# Perl binding of libpedk.so, the B200-native k-mer hot path of PED.
# The subs mirror the ped.pl subs they replace (see INTEGRATION.md):
#   PedK::mk_sort_uniq($wd, $subject, $clipping)   -> ped.pl mkSortUniq  (1365-1410); returns (clipping value, automatic?)
#   PedK::count_kmer($wd, $subject, $k)            -> ped.pl countKmer   (784-815)
#   PedK::lbc_kmer($wd, $subject, $k)              -> ped.pl lbcKmer     (817-831)
#   PedK::join_kmer($wd, $t, $c, $k, $tmpdir, $ref)-> ped.pl joinKmer    (687-701) + 374-378
#   PedK::mk_control_read($wd, $ref, $file)        -> ped.pl mkChrFromFile + mkControlRead (1027-1136)
#   PedK::mk_ref20_uniq($wd, $ref, $file, $k)      -> ped.pl mk20        (1213-1249: mk20mer + mkUniq)
#   PedK::map_kmer($wd, $t, $ref, $tmpdir)         -> ped.pl mapKmer     (625-685)
#   PedK::verify_kmer($wd, $t, $c, $ref, $tmpdir)  -> ped.pl snpMkT, sortSeq, joinTarget, snpMkC, sortSeq, joinControl,
#                                                     kmerReadCount (2087-2305, 2576-2641, 2388-2500)
#   PedK::to_vcf($wd, $t)                          -> ped.pl toVcf, method kmer (1532-1574)
#   PedK::cnv($wd, $t, $c, $ref, $k)               -> ped.pl cnvJoin     (934-989)
# Every sub dies with the library's message on failure; there is no CPU fallback.
use strict;
use warnings;
our $VERSION = '0.01';
require XSLoader;
XSLoader::load('PedK', $VERSION);
our @STAGES = qw(ingest dedup extract sort count expand rowsort edges copy exchange);
1;
