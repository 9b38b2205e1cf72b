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
# `ref=REF,file=REF.fa` (no target) builds the control reads of a reference like
# ped.pl's mkRef does (ped.pl:1027-1136: <wd>/REF/chrNAME and REF/sort_uniq/); a
# later `target=T,control=REF` then compares T with the tiled reference.
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
    PedK::mk_control_read($wd, $opt{ref}, $opt{file});
    PedK::close_device();
    print "control reads of $opt{ref} written to $wd/$opt{ref}/sort_uniq\n";
    exit 0;
}
die "ped_kmer.pl runs the k-mer comparison itself; with ref= the mapping stages of ped.pl follow: use ped.pl with the PedK patch of INTEGRATION.md\n"
    if defined $opt{ref} && $opt{ref} ne "";
die "method must be kmer\n" if defined $opt{method} && $opt{method} ne "kmer";
my ($target, $control) = ($opt{target}, $opt{control});
die "target= and control= are required\n" unless defined $target && $target ne "" && defined $control && $control ne "";
my $wd = (defined $opt{wd} && $opt{wd} ne "") ? $opt{wd} : getcwd();
my $tmpdir = (defined $opt{tmpdir} && $opt{tmpdir} ne "") ? $opt{tmpdir} : "$wd/$target/tmp";
my $clipping = (defined $opt{clipping} && $opt{clipping} ne "") ? int($opt{clipping}) : 0;
my $k = int($opt{k});

my $start = time;
my $start_stamp = localtime;
open(my $REPORT, ">", "$wd/$target/$target.report") or die "cannot write $wd/$target/$target.report: $!\n";
open(my $LOG, ">", "$wd/$target/$target.log") or die "cannot write $wd/$target/$target.log: $!\n";
print $LOG $log, "method : kmer\n";

sub report {
    my $m = shift;
    my @t = localtime;
    my $now = sprintf("%04d-%02d-%02d %02d:%02d:%02d", $t[5] + 1900, $t[4] + 1, @t[3, 2, 1, 0]);
    $m = "$now : $target : $m\n";
    print $m;
    print $REPORT $m;
}

report("Job started: kmer method (B200)");
PedK::open_device();
foreach my $s ($target, $control) {
    next if -e "$wd/$s/sort_uniq/$s.sort_uniq.TTT.gz";
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
foreach my $s ($target, $control) {
    next if -e "$wd/$s/lbc/$s.lbc.TTT.gz";
    report("Joining k-mers for $s");
    PedK::lbc_kmer($wd, $s, $k);
}
report("Detecting SNPs between $target and $control");
PedK::join_kmer($wd, $target, $control, $k, $tmpdir, 0);
my $ms = PedK::stage_ms();
report("GPU stage times (ms): " . join(" ", map { "$PedK::STAGES[$_]=" . sprintf("%.1f", $ms->[$_]) } 0 .. $#PedK::STAGES));
PedK::close_device();
my $elapsed = time - $start;
report("Job finished: kmer method.\n$elapsed seconds elapsed.");
print $LOG "start : $start_stamp\nend : " . localtime() . "\nseconds : $elapsed\n";
close($LOG);
close($REPORT);
