/* PedK.xs -- XS glue between ped.pl and libpedk.so (include/pedk.h).
 *
 * FFI::Platypus is not available in the target image (SURVEY.md 8b), XS is.
 * One sub per ped.pl sub that the GPU library replaces; each croaks with the
 * library's message on failure so that ped.pl dies instead of leaving a
 * partial stage behind.
 */
#define PERL_NO_GET_CONTEXT
#include "EXTERN.h"
#include "perl.h"
#include "XSUB.h"

#include "pedk.h"

static pedk_ctx *g_ctx = NULL;

static pedk_ctx *ctx_or_croak(pTHX) {
    if (!g_ctx) {
        int dev = 0;
        const char *e = getenv("PEDK_DEVICE");
        if (e) dev = atoi(e);
        int rc = pedk_open(&g_ctx, &dev, 1);
        if (rc != PEDK_OK) croak("PedK: %s (%s)", pedk_last_error(NULL), pedk_strerror(rc));
    }
    return g_ctx;
}

static void check(pTHX_ int rc, const char *what) {
    if (rc != PEDK_OK) croak("PedK::%s: %s (%s)", what, pedk_last_error(g_ctx), pedk_strerror(rc));
}

MODULE = PedK    PACKAGE = PedK

PROTOTYPES: DISABLE

int
abi_version()
  CODE:
    RETVAL = pedk_abi_version();
  OUTPUT:
    RETVAL

void
open_device()
  CODE:
    ctx_or_croak(aTHX);

void
close_device()
  CODE:
    if (g_ctx) { pedk_close(g_ctx); g_ctx = NULL; }

void
mk_sort_uniq(wd, subject, clipping)
    const char *wd
    const char *subject
    int clipping
  PPCODE:
    {
        /* returns (clipping value for the log line, 1 if it was chosen automatically) */
        int used = 0, automatic = 0;
        check(aTHX_ pedk_mk_sort_uniq(ctx_or_croak(aTHX), wd, subject, clipping, &used, &automatic), "mk_sort_uniq");
        EXTEND(SP, 2);
        PUSHs(sv_2mortal(newSViv(used)));
        PUSHs(sv_2mortal(newSViv(automatic)));
    }

void
mk_control_read(wd, ref, file)
    const char *wd
    const char *ref
    const char *file
  CODE:
    /* mkChrFromFile + mkControlRead (ped.pl:1027-1136) */
    check(aTHX_ pedk_mk_control_read(ctx_or_croak(aTHX), wd, ref, file), "mk_control_read");

void
count_kmer(wd, subject, k)
    const char *wd
    const char *subject
    int k
  CODE:
    check(aTHX_ pedk_count_kmer(ctx_or_croak(aTHX), wd, subject, k), "count_kmer");

void
lbc_kmer(wd, subject, k)
    const char *wd
    const char *subject
    int k
  CODE:
    check(aTHX_ pedk_lbc_kmer(ctx_or_croak(aTHX), wd, subject, k), "lbc_kmer");

void
join_kmer(wd, target, control, k, tmpdir, have_ref)
    const char *wd
    const char *target
    const char *control
    int k
    const char *tmpdir
    int have_ref
  CODE:
    check(aTHX_ pedk_join_kmer(ctx_or_croak(aTHX), wd, target, control, k, tmpdir, have_ref), "join_kmer");

void
mk_ref20_uniq(wd, ref, file, k)
    const char *wd
    const char *ref
    SV *file
    int k
  CODE:
    /* mk20 = mk20mer + mkUniq (ped.pl:1138-1249); file may be undef when the reference is still resident */
    check(aTHX_ pedk_mk_ref20_uniq(ctx_or_croak(aTHX), wd, ref, SvOK(file) ? SvPV_nolen(file) : NULL, k), "mk_ref20_uniq");

void
map_kmer(wd, target, ref, tmpdir)
    const char *wd
    const char *target
    const char *ref
    const char *tmpdir
  CODE:
    /* mapKmer + mapKmerSub (ped.pl:625-685) */
    check(aTHX_ pedk_map_kmer(ctx_or_croak(aTHX), wd, target, ref, tmpdir), "map_kmer");

void
verify_kmer(wd, target, control, ref, tmpdir)
    const char *wd
    const char *target
    const char *control
    const char *ref
    const char *tmpdir
  CODE:
    /* snpMkT, sortSeq, joinTarget, snpMkC, sortSeq, joinControl, kmerReadCount (ped.pl:380-386) */
    check(aTHX_ pedk_verify_kmer(ctx_or_croak(aTHX), wd, target, control, ref, tmpdir), "verify_kmer");

void
to_vcf(wd, target)
    const char *wd
    const char *target
  CODE:
    /* toVcf, method kmer (ped.pl:387, 1532-1574) */
    check(aTHX_ pedk_to_vcf(ctx_or_croak(aTHX), wd, target), "to_vcf");

void
cnv(wd, target, control, ref, k)
    const char *wd
    const char *target
    const char *control
    const char *ref
    int k
  CODE:
    /* cnvJoin + cnvJoinSub (ped.pl:934-989) */
    check(aTHX_ pedk_cnv(ctx_or_croak(aTHX), wd, target, control, ref, k), "cnv");

SV *
stage_ms()
  CODE:
    {
        pedk_stats st;
        AV *av = newAV();
        int i;
        check(aTHX_ pedk_get_stats(ctx_or_croak(aTHX), &st), "stage_ms");
        for (i = 0; i < PEDK_N_STAGES; ++i) av_push(av, newSVnv(st.stage_ms[i]));
        RETVAL = newRV_noinc((SV *)av);
    }
  OUTPUT:
    RETVAL
