// cs_allele.cu — allele (coverage + major-allele-fraction) variant of the sampler.
#include "cs_common.cuh"

extern "C" int cs_ncrp_gibbs_allele(int N, int R, const double *Xcov, const double *Xallele, int K0,
                                    const int *U0, const int *Z0, const double *sig0_cov,
                                    const double *sig0_allele, double alpha, double beta, int niter,
                                    uint32_t seed, int maxcp, const cs_gibbs_opts *opts, int *Zall,
                                    double *sig_cov_all, double *sig_allele_all, double *LL, double *L0,
                                    int *kt_all, int64_t ucap_rows, int64_t *u_off, int *u_labels, int *u_state,
                                    double *u_cov, double *u_allele)
{
    cs_set_error("cs_ncrp_gibbs_allele: not built yet");
    return CS_ERR_UNSUPPORTED;
}
