#!/usr/bin/perl
#
# ped_kmer.pl -- the k-mer method of PED with its hot path on a B200.
#
# Same command line as ped.pl (reference ped.pl:99-127): one comma-joined
# argument `target=T,control=C[,wd=DIR][,tmpdir=DIR][,clipping=L][,k=K]`, the
# positional form `T C` and the environment form.  Same files as ped.pl's
# method=kmer without ref= (ped.pl:368-378): <wd>/<s>/sort_uniq/, count/, lbc/,
# <wd>/<t>/<t>.kmer.snp, <t>.log with the `clipping :` lines, <t>.report; a stage
# is skipped when its TTT file exists (ped.pl:291-297, 369-372).  The heavy
# subs are calls into libpedk.so through PedK.xs; nothing here forks.
#
# `ref=REF,file=REF.fa` (no target) builds a reference like ped.pl's mkRef does
# (ped.pl:1017-1025: <wd>/REF/chrNAME, REF/ref20_uniq.TAG.gz and the control reads in
# REF/sort_uniq/); a later `target=T,control=REF` then compares T with the tiled reference.
# With `ref=REF` the per-tag edge files go to tmpdir and are mapped to positions
# (<tmpdir>/<t>.map.TAG, ped.pl:625-685), the candidates are verified against the reads of both
# samples and genotyped (<wd>/<t>/<t>.kmer.snp with 20 columns, ped.pl:380-386) and written as VCF
# (<t>.kmer.vcf, ped.pl:387); `primer` (ped.pl:388) stays ped.pl's.
# `method=cnv` writes <wd>/<t>/<t>.<c>.cnv (ped.pl:362-366, 934-989).
#
use strict;
use warnings;
use FindBin;
use lib "$FindBin::Bin/lib";
use PedK;
use Cwd qw(getcwd);

$ENV{LANG} = "C";
my %opt = (k => 20);
my $arg0 = defined $ARGV[0] ? $ARGV[0] : "";
my $log = "script : ped_kmer.pl\nargument : $arg0\n";
if ($arg0 =~ /target/ || $arg0 =~ /ref=.*file=|file=.*ref=/) {
    foreach (sort split(',', $arg0)) {
        next if $_ eq "";
        my ($name, $val) = split('=', $_);
        $opt{$name} = $val;
        $log .= "$name : " . (defined $val ? $val : "") . "\n";
    }
} elsif (defined $ARGV[1] && $ARGV[1] ne "") {
    @opt{qw(target control)} = @ARGV[0, 1];
} elsif (defined $ENV{target} && $ENV{target} ne "") {
    $opt{$_} = $ENV{$_} foreach grep { defined $ENV{$_} } qw(target control wd tmpdir clipping);
} else {
    die "usage: perl ped_kmer.pl target=T,control=C[,wd=DIR][,clipping=L]\n";
}
if (defined $opt{ref} && $opt{ref} ne "" && defined $opt{file} && $opt{file} ne "" && !defined $opt{target}) {
    # mkRef for a local FASTA (ped.pl:1017-1025): chromosome files + control reads
    my $wd = (defined $opt{wd} && $opt{wd} ne "") ? $opt{wd} : getcwd();
    PedK::open_device();
    PedK::mk_control_read($wd, $opt{ref}, $opt{file});     # mkChrFromFile + mkControlRead
    PedK::mk_ref20_uniq($wd, $opt{ref}, undef, 20);        # mk20 (the reference is still resident)
    PedK::close_device();
    print "reference $opt{ref}: chromosome files, ref20_uniq and control reads written to $wd/$opt{ref}\n";
    exit 0;
}
my $ref = (defined $opt{ref} && $opt{ref} ne "") ? $opt{ref} : "";
my $method = (defined $opt{method} && $opt{method} ne "") ? $opt{method} : "kmer";
die "method must be kmer or cnv\n" if $method ne "kmer" && $method ne "cnv";
die "method=cnv needs ref=\n" if $method eq "cnv" && $ref eq "";
my ($target, $control) = ($opt{target}, $opt{control});
$control = $ref if (!defined $control || $control eq "") && $ref ne "";    # ped.pl:193
die "target= and control= are required\n" unless defined $target && $target ne "" && defined $control && $control ne "";
my $wd = (defined $opt{wd} && $opt{wd} ne "") ? $opt{wd} : getcwd();
my $tmpdir = (defined $opt{tmpdir} && $opt{tmpdir} ne "") ? $opt{tmpdir} : "$wd/$target/tmp";
my $clipping = (defined $opt{clipping} && $opt{clipping} ne "") ? int($opt{clipping}) : 0;
my $k = int($opt{k});

my $start = time;
my $start_stamp = localtime;
open(my $REPORT, ">", "$wd/$target/$target.report") or die "cannot write $wd/$target/$target.report: $!\n";
open(my $LOG, ">", "$wd/$target/$target.log") or die "cannot write $wd/$target/$target.log: $!\n";
print $LOG $log, "method : $method\n";

sub report {
    my $m = shift;
    my @t = localtime;
    my $now = sprintf("%04d-%02d-%02d %02d:%02d:%02d", $t[5] + 1900, $t[4] + 1, @t[3, 2, 1, 0]);
    $m = "$now : $target : $m\n";
    print $m;
    print $REPORT $m;
}

report("Job started: $method method (B200)");
PedK::open_device();
foreach my $s ($target, $control) {
    next if -e "$wd/$s/sort_uniq/$s.sort_uniq.TTT.gz";
    next if $s eq $ref;    # the reference's control reads come from `ref=,file=` (ped.pl:295)
    report("Generating sort_uniq files for $s");
    my ($used, $automatic) = PedK::mk_sort_uniq($wd, $s, $clipping);
    print $LOG "clipping : $used\n";
    $clipping = $used if $automatic;    # ped.pl:1483 assigns its global: it sticks for the control
}
foreach my $s ($target, $control) {
    next if -e "$wd/$s/count/$s.count.TTT.gz";
    report("Generating k-mers for $s");
    PedK::count_kmer($wd, $s, $k);
}
if ($method eq "cnv") {
    report("Joining k-mers between $target and $control (cnv)");
    PedK::cnv($wd, $target, $control, $ref, $k);
} else {
    foreach my $s ($target, $control) {
        next if -e "$wd/$s/lbc/$s.lbc.TTT.gz";
        report("Joining k-mers for $s");
        PedK::lbc_kmer($wd, $s, $k);
    }
    report("Detecting SNPs between $target and $control");
    mkdir $tmpdir if $ref ne "" && !-e $tmpdir;
    PedK::join_kmer($wd, $target, $control, $k, $tmpdir, $ref ne "" ? 1 : 0);
    if ($ref ne "") {
        report("Mapping SNPs between $target and $control");
        PedK::map_kmer($wd, $target, $ref, $tmpdir);
        report("Verifying SNPs with the reads of $target and $control");   # ped.pl:380-386
        PedK::verify_kmer($wd, $target, $control, $ref, $tmpdir);
        report("Converting to VCF format");                                # ped.pl:387
        PedK::to_vcf($wd, $target);
    }
}
my $ms = PedK::stage_ms();
report("GPU stage times (ms): " . join(" ", map { "$PedK::STAGES[$_]=" . sprintf("%.1f", $ms->[$_]) } 0 .. $#PedK::STAGES));
PedK::close_device();
my $elapsed = time - $start;
report("Job finished: $method method.\n$elapsed seconds elapsed.");
print $LOG "start : $start_stamp\nend : " . localtime() . "\nseconds : $elapsed\n";
close($LOG);
close($REPORT);
