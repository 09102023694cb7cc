// TMA-staged SpMV variant (placeholder until the kernel lands; the MARCH variant is the product path).
#include "common.cuh"

int s3d_spmv_tma(s3d_ctx *ctx, const s3d_grid *, const double *, const double *, double *, const double *,
                 const double *, bool, double *, const int *)
{
	return s3d_fail(ctx, S3D_ERR_ARG, "S3D_SPMV_TMA is not built in this revision");
}
