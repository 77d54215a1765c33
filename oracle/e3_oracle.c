/*
 * e3_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or
 * executed from the product path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it).
 *
 * Plain-C CPU restatement of the arithmetic that elektronn3's UNet hot path
 * delegates to torch (the algorithm lives in PyTorch ATen/oneDNN, unpinned
 * `torch>=1.6.0`, reference requirements.txt:1).  Each function restates the
 * published semantics of one torch operator at the call site named beside it.
 * Layout: contiguous NCDHW fp32 exactly as the reference uses (no
 * channels_last anywhere in the reference).  Sums are accumulated in double so
 * that the oracle is the tighter arbiter between the fp32 CPU reference and
 * the TF32 GPU path.
 *
 * Parity pin: the reference has NO golden vectors of its own for this path
 * (its only tests assert output shape, models/unet.py:938-1026), so this
 * oracle is pinned against outputs of the reference itself, generated in the
 * build container by oracle/gen_golden.py and committed under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define IDX5(n, c, d, h, w, C, D, H, W) \
    (((((int64_t)(n) * (C) + (c)) * (D) + (d)) * (H) + (h)) * (W) + (w))

/* ------------------------------------------------------------------------
 * nn.Conv3d cross-correlation, stride 1, zero padding (pd,ph,pw).
 * Call sites: conv3 models/unet.py:131-149 (k=3 or planar (1,3,3), pad 1 /
 * (0,1,1) / 0), conv1 models/unet.py:178-180 (k=1).
 * w: (Co,Ci,kd,kh,kw); y: (N,Co,Do,Ho,Wo) with Do = D + 2pd - kd + 1.
 * ---------------------------------------------------------------------- */
void e3o_conv3d_fwd(const float *x, const float *w, const float *b, float *y,
                    int N, int Ci, int D, int H, int W, int Co,
                    int kd, int kh, int kw, int pd, int ph, int pw)
{
    int Do = D + 2 * pd - kd + 1, Ho = H + 2 * ph - kh + 1, Wo = W + 2 * pw - kw + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int co = 0; co < Co; co++) {
            double *acc = (double *)malloc(sizeof(double) * Wo);
            for (int d = 0; d < Do; d++)
                for (int h = 0; h < Ho; h++) {
                    for (int x0 = 0; x0 < Wo; x0++) acc[x0] = b ? (double)b[co] : 0.0;
                    for (int ci = 0; ci < Ci; ci++)
                        for (int i = 0; i < kd; i++) {
                            int zd = d + i - pd;
                            if (zd < 0 || zd >= D) continue;
                            for (int j = 0; j < kh; j++) {
                                int zh = h + j - ph;
                                if (zh < 0 || zh >= H) continue;
                                const float *xr = x + IDX5(n, ci, zd, zh, 0, Ci, D, H, W);
                                for (int k = 0; k < kw; k++) {
                                    double wv = w[((((int64_t)co * Ci + ci) * kd + i) * kh + j) * kw + k];
                                    int lo = pw - k; if (lo < 0) lo = 0;
                                    int hi = W + pw - k; if (hi > Wo) hi = Wo;
                                    for (int x0 = lo; x0 < hi; x0++)
                                        acc[x0] += wv * (double)xr[x0 + k - pw];
                                }
                            }
                        }
                    float *yr = y + IDX5(n, co, d, h, 0, Co, Do, Ho, Wo);
                    for (int x0 = 0; x0 < Wo; x0++) yr[x0] = (float)acc[x0];
                }
            free(acc);
        }
}

/* Backward of the above (what torch autograd computes):
 * dx = full correlation of dy with flipped w; dw = sum over voxels; db = sum dy.
 * Any of dx/dw/db may be NULL. */
void e3o_conv3d_bwd(const float *x, const float *w, const float *dy,
                    float *dx, float *dw, float *db,
                    int N, int Ci, int D, int H, int W, int Co,
                    int kd, int kh, int kw, int pd, int ph, int pw)
{
    int Do = D + 2 * pd - kd + 1, Ho = H + 2 * ph - kh + 1, Wo = W + 2 * pw - kw + 1;
    if (dx) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int n = 0; n < N; n++)
            for (int ci = 0; ci < Ci; ci++) {
                double *acc = (double *)malloc(sizeof(double) * W);
                for (int d = 0; d < D; d++)
                    for (int h = 0; h < H; h++) {
                        for (int x0 = 0; x0 < W; x0++) acc[x0] = 0.0;
                        for (int co = 0; co < Co; co++)
                            for (int i = 0; i < kd; i++) {
                                int od = d - i + pd;
                                if (od < 0 || od >= Do) continue;
                                for (int j = 0; j < kh; j++) {
                                    int oh = h - j + ph;
                                    if (oh < 0 || oh >= Ho) continue;
                                    const float *dr = dy + IDX5(n, co, od, oh, 0, Co, Do, Ho, Wo);
                                    for (int k = 0; k < kw; k++) {
                                        double wv = w[((((int64_t)co * Ci + ci) * kd + i) * kh + j) * kw + k];
                                        /* ow = x0 - k + pw in [0,Wo) */
                                        int lo = k - pw; if (lo < 0) lo = 0;
                                        int hi = Wo + k - pw; if (hi > W) hi = W;
                                        for (int x0 = lo; x0 < hi; x0++)
                                            acc[x0] += wv * (double)dr[x0 - k + pw];
                                    }
                                }
                            }
                        float *xr = dx + IDX5(n, ci, d, h, 0, Ci, D, H, W);
                        for (int x0 = 0; x0 < W; x0++) xr[x0] = (float)acc[x0];
                    }
                free(acc);
            }
    }
    if (dw) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int co = 0; co < Co; co++)
            for (int ci = 0; ci < Ci; ci++)
                for (int i = 0; i < kd; i++)
                    for (int j = 0; j < kh; j++)
                        for (int k = 0; k < kw; k++) {
                            double s = 0.0;
                            for (int n = 0; n < N; n++)
                                for (int d = 0; d < Do; d++) {
                                    int zd = d + i - pd;
                                    if (zd < 0 || zd >= D) continue;
                                    for (int h = 0; h < Ho; h++) {
                                        int zh = h + j - ph;
                                        if (zh < 0 || zh >= H) continue;
                                        const float *xr = x + IDX5(n, ci, zd, zh, 0, Ci, D, H, W);
                                        const float *dr = dy + IDX5(n, co, d, h, 0, Co, Do, Ho, Wo);
                                        int lo = pw - k; if (lo < 0) lo = 0;
                                        int hi = W + pw - k; if (hi > Wo) hi = Wo;
                                        for (int x0 = lo; x0 < hi; x0++)
                                            s += (double)dr[x0] * (double)xr[x0 + k - pw];
                                    }
                                }
                            dw[((((int64_t)co * Ci + ci) * kd + i) * kh + j) * kw + k] = (float)s;
                        }
    }
    if (db) {
#pragma omp parallel for schedule(static)
        for (int co = 0; co < Co; co++) {
            double s = 0.0;
            for (int n = 0; n < N; n++) {
                const float *dr = dy + IDX5(n, co, 0, 0, 0, Co, Do, Ho, Wo);
                for (int64_t v = 0; v < (int64_t)Do * Ho * Wo; v++) s += dr[v];
            }
            db[co] = (float)s;
        }
    }
}

/* ------------------------------------------------------------------------
 * nn.ConvTranspose3d with kernel == stride == (sd,sh,sw) (2,2,2 or planar
 * (1,2,2)); upconv2 models/unet.py:152-165.  w: (Ci,Co,sd,sh,sw).
 * y[n,co,sd*d+i,sh*h+j,sw*w+k] = b[co] + sum_ci x[n,ci,d,h,w] w[ci,co,i,j,k]
 * ---------------------------------------------------------------------- */
void e3o_convT_fwd(const float *x, const float *w, const float *b, float *y,
                   int N, int Ci, int D, int H, int W, int Co, int sd, int sh, int sw)
{
    int Do = D * sd, Ho = H * sh, Wo = W * sw;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int co = 0; co < Co; co++)
            for (int d = 0; d < D; d++)
                for (int h = 0; h < H; h++)
                    for (int x0 = 0; x0 < W; x0++)
                        for (int i = 0; i < sd; i++)
                            for (int j = 0; j < sh; j++)
                                for (int k = 0; k < sw; k++) {
                                    double s = b ? (double)b[co] : 0.0;
                                    for (int ci = 0; ci < Ci; ci++)
                                        s += (double)x[IDX5(n, ci, d, h, x0, Ci, D, H, W)] *
                                             (double)w[((((int64_t)ci * Co + co) * sd + i) * sh + j) * sw + k];
                                    y[IDX5(n, co, d * sd + i, h * sh + j, x0 * sw + k, Co, Do, Ho, Wo)] = (float)s;
                                }
}

void e3o_convT_bwd(const float *x, const float *w, const float *dy,
                   float *dx, float *dw, float *db,
                   int N, int Ci, int D, int H, int W, int Co, int sd, int sh, int sw)
{
    int Do = D * sd, Ho = H * sh, Wo = W * sw;
    if (dx) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int n = 0; n < N; n++)
            for (int ci = 0; ci < Ci; ci++)
                for (int d = 0; d < D; d++)
                    for (int h = 0; h < H; h++)
                        for (int x0 = 0; x0 < W; x0++) {
                            double s = 0.0;
                            for (int co = 0; co < Co; co++)
                                for (int i = 0; i < sd; i++)
                                    for (int j = 0; j < sh; j++)
                                        for (int k = 0; k < sw; k++)
                                            s += (double)dy[IDX5(n, co, d * sd + i, h * sh + j, x0 * sw + k, Co, Do, Ho, Wo)] *
                                                 (double)w[((((int64_t)ci * Co + co) * sd + i) * sh + j) * sw + k];
                            dx[IDX5(n, ci, d, h, x0, Ci, D, H, W)] = (float)s;
                        }
    }
    if (dw) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int ci = 0; ci < Ci; ci++)
            for (int co = 0; co < Co; co++)
                for (int i = 0; i < sd; i++)
                    for (int j = 0; j < sh; j++)
                        for (int k = 0; k < sw; k++) {
                            double s = 0.0;
                            for (int n = 0; n < N; n++)
                                for (int d = 0; d < D; d++)
                                    for (int h = 0; h < H; h++)
                                        for (int x0 = 0; x0 < W; x0++)
                                            s += (double)x[IDX5(n, ci, d, h, x0, Ci, D, H, W)] *
                                                 (double)dy[IDX5(n, co, d * sd + i, h * sh + j, x0 * sw + k, Co, Do, Ho, Wo)];
                            dw[((((int64_t)ci * Co + co) * sd + i) * sh + j) * sw + k] = (float)s;
                        }
    }
    if (db) {
        for (int co = 0; co < Co; co++) {
            double s = 0.0;
            for (int n = 0; n < N; n++) {
                const float *dr = dy + IDX5(n, co, 0, 0, 0, Co, Do, Ho, Wo);
                for (int64_t v = 0; v < (int64_t)Do * Ho * Wo; v++) s += dr[v];
            }
            db[co] = (float)s;
        }
    }
}

/* ------------------------------------------------------------------------
 * nn.GroupNorm(G, C), eps 1e-5, affine; get_normalization
 * models/unet.py:81-91.  Stats per (n, g) over C/G x S voxels, biased var.
 * 'instance' (unet.py:92-98, affine=False there) == G = C with gamma=1,beta=0.
 * mean/rstd: (N,G) outputs (may be NULL).
 * ---------------------------------------------------------------------- */
void e3o_groupnorm_fwd(const float *x, const float *gamma, const float *beta, float *y,
                       float *mean, float *rstd, int N, int C, int64_t S, int G, float eps)
{
    int cg = C / G;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int g = 0; g < G; g++) {
            const float *xp = x + ((int64_t)n * C + (int64_t)g * cg) * S;
            int64_t cnt = (int64_t)cg * S;
            double s = 0.0, ss = 0.0;
            for (int64_t i = 0; i < cnt; i++) s += xp[i];
            double mu = s / (double)cnt;
            for (int64_t i = 0; i < cnt; i++) { double t = xp[i] - mu; ss += t * t; }
            double r = 1.0 / sqrt(ss / (double)cnt + (double)eps);
            if (mean) mean[n * G + g] = (float)mu;
            if (rstd) rstd[n * G + g] = (float)r;
            for (int c = 0; c < cg; c++) {
                double ga = gamma ? gamma[g * cg + c] : 1.0, be = beta ? beta[g * cg + c] : 0.0;
                float *yp = y + ((int64_t)n * C + (int64_t)g * cg + c) * S;
                const float *xc = xp + (int64_t)c * S;
                for (int64_t i = 0; i < S; i++) yp[i] = (float)((xc[i] - mu) * r * ga + be);
            }
        }
}

/* dx = r (g - mean(g) - xhat mean(g xhat)), g = dy*gamma; dgamma = sum dy xhat; dbeta = sum dy */
void e3o_groupnorm_bwd(const float *x, const float *gamma, const float *dy,
                       float *dx, float *dgamma, float *dbeta,
                       int N, int C, int64_t S, int G, float eps)
{
    int cg = C / G;
    double *dga = (double *)calloc((size_t)N * C, sizeof(double));
    double *dbe = (double *)calloc((size_t)N * C, sizeof(double));
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int g = 0; g < G; g++) {
            const float *xp = x + ((int64_t)n * C + (int64_t)g * cg) * S;
            const float *dp = dy + ((int64_t)n * C + (int64_t)g * cg) * S;
            int64_t cnt = (int64_t)cg * S;
            double s = 0.0, ss = 0.0;
            for (int64_t i = 0; i < cnt; i++) s += xp[i];
            double mu = s / (double)cnt;
            for (int64_t i = 0; i < cnt; i++) { double t = xp[i] - mu; ss += t * t; }
            double r = 1.0 / sqrt(ss / (double)cnt + (double)eps);
            double m1 = 0.0, m2 = 0.0;
            for (int c = 0; c < cg; c++) {
                double ga = gamma ? gamma[g * cg + c] : 1.0;
                double a1 = 0.0, a2 = 0.0;
                for (int64_t i = 0; i < S; i++) {
                    double xh = (xp[c * S + i] - mu) * r, d = dp[c * S + i];
                    a1 += d; a2 += d * xh;
                }
                dbe[n * C + g * cg + c] = a1; dga[n * C + g * cg + c] = a2;
                m1 += ga * a1; m2 += ga * a2;
            }
            m1 /= (double)cnt; m2 /= (double)cnt;
            if (dx)
                for (int c = 0; c < cg; c++) {
                    double ga = gamma ? gamma[g * cg + c] : 1.0;
                    for (int64_t i = 0; i < S; i++) {
                        double xh = (xp[c * S + i] - mu) * r;
                        dx[((int64_t)n * C + (int64_t)g * cg + c) * S + i] =
                            (float)(r * (dp[c * S + i] * ga - m1 - xh * m2));
                    }
                }
        }
    for (int c = 0; c < C; c++) {
        double a = 0.0, b2 = 0.0;
        for (int n = 0; n < N; n++) { a += dga[n * C + c]; b2 += dbe[n * C + c]; }
        if (dgamma) dgamma[c] = (float)a;
        if (dbeta) dbeta[c] = (float)b2;
    }
    free(dga); free(dbe);
}

/* ------------------------------------------------------------------------
 * nn.BatchNorm3d/2d (unet.py:99-105), eps 1e-5, momentum 0.1.
 * training=1: batch stats (biased var normalises, unbiased var goes into
 * running_var), running stats updated in place; training=0: running stats.
 * save_mean/save_rstd (C) optional.
 * ---------------------------------------------------------------------- */
void e3o_batchnorm_fwd(const float *x, const float *gamma, const float *beta,
                       float *running_mean, float *running_var, float *y,
                       float *save_mean, float *save_rstd,
                       int N, int C, int64_t S, float eps, float momentum, int training)
{
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; c++) {
        double mu, var;
        int64_t cnt = (int64_t)N * S;
        if (training) {
            double s = 0.0, ss = 0.0;
            for (int n = 0; n < N; n++) {
                const float *xp = x + ((int64_t)n * C + c) * S;
                for (int64_t i = 0; i < S; i++) s += xp[i];
            }
            mu = s / (double)cnt;
            for (int n = 0; n < N; n++) {
                const float *xp = x + ((int64_t)n * C + c) * S;
                for (int64_t i = 0; i < S; i++) { double t = xp[i] - mu; ss += t * t; }
            }
            var = ss / (double)cnt;
            if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mu);
            if (running_var) {
                double unb = cnt > 1 ? ss / (double)(cnt - 1) : var;
                running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
            }
        } else { mu = running_mean[c]; var = running_var[c]; }
        double r = 1.0 / sqrt(var + (double)eps);
        if (save_mean) save_mean[c] = (float)mu;
        if (save_rstd) save_rstd[c] = (float)r;
        double ga = gamma ? gamma[c] : 1.0, be = beta ? beta[c] : 0.0;
        for (int n = 0; n < N; n++) {
            const float *xp = x + ((int64_t)n * C + c) * S;
            float *yp = y + ((int64_t)n * C + c) * S;
            for (int64_t i = 0; i < S; i++) yp[i] = (float)((xp[i] - mu) * r * ga + be);
        }
    }
}

/* training-mode backward (batch statistics take part in the graph) */
void e3o_batchnorm_bwd(const float *x, const float *gamma, const float *dy,
                       float *dx, float *dgamma, float *dbeta,
                       int N, int C, int64_t S, float eps)
{
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; c++) {
        int64_t cnt = (int64_t)N * S;
        double s = 0.0, ss = 0.0;
        for (int n = 0; n < N; n++) {
            const float *xp = x + ((int64_t)n * C + c) * S;
            for (int64_t i = 0; i < S; i++) s += xp[i];
        }
        double mu = s / (double)cnt;
        for (int n = 0; n < N; n++) {
            const float *xp = x + ((int64_t)n * C + c) * S;
            for (int64_t i = 0; i < S; i++) { double t = xp[i] - mu; ss += t * t; }
        }
        double r = 1.0 / sqrt(ss / (double)cnt + (double)eps);
        double a1 = 0.0, a2 = 0.0;
        for (int n = 0; n < N; n++) {
            const float *xp = x + ((int64_t)n * C + c) * S;
            const float *dp = dy + ((int64_t)n * C + c) * S;
            for (int64_t i = 0; i < S; i++) { a1 += dp[i]; a2 += dp[i] * (xp[i] - mu) * r; }
        }
        if (dgamma) dgamma[c] = (float)a2;
        if (dbeta) dbeta[c] = (float)a1;
        double ga = gamma ? gamma[c] : 1.0;
        double m1 = ga * a1 / (double)cnt, m2 = ga * a2 / (double)cnt;
        if (dx)
            for (int n = 0; n < N; n++) {
                const float *xp = x + ((int64_t)n * C + c) * S;
                const float *dp = dy + ((int64_t)n * C + c) * S;
                float *op = dx + ((int64_t)n * C + c) * S;
                for (int64_t i = 0; i < S; i++)
                    op[i] = (float)(r * (dp[i] * ga - m1 - (xp[i] - mu) * r * m2));
            }
    }
}

/* nn.ReLU, get_activation models/unet.py:183-186 */
void e3o_relu_fwd(const float *x, float *y, int64_t n)
{
    for (int64_t i = 0; i < n; i++) y[i] = x[i] > 0.f ? x[i] : 0.f;
}
void e3o_relu_bwd(const float *y, const float *dy, float *dx, int64_t n)
{
    for (int64_t i = 0; i < n; i++) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

/* ------------------------------------------------------------------------
 * nn.MaxPool3d(kernel=(kd,kh,kw), stride=kernel, ceil_mode=True);
 * DownConv models/unet.py:225-229.  Out extent = ceil(D/kd); the last window
 * is partial for odd extents.  idx (optional) = flat index into the (D,H,W)
 * plane of the first maximum in (d,h,w) scan order (torch CPU semantics).
 * ---------------------------------------------------------------------- */
void e3o_maxpool_fwd(const float *x, float *y, int64_t *idx,
                     int N, int C, int D, int H, int W, int kd, int kh, int kw)
{
    int Do = (D + kd - 1) / kd, Ho = (H + kh - 1) / kh, Wo = (W + kw - 1) / kw;
#pragma omp parallel for schedule(static)
    for (int64_t nc = 0; nc < (int64_t)N * C; nc++) {
        const float *xp = x + nc * D * H * W;
        for (int d = 0; d < Do; d++)
            for (int h = 0; h < Ho; h++)
                for (int w0 = 0; w0 < Wo; w0++) {
                    float best = -FLT_MAX; int64_t bi = -1;
                    for (int i = 0; i < kd; i++)
                        for (int j = 0; j < kh; j++)
                            for (int k = 0; k < kw; k++) {
                                int zd = d * kd + i, zh = h * kh + j, zw = w0 * kw + k;
                                if (zd >= D || zh >= H || zw >= W) continue;
                                int64_t ii = ((int64_t)zd * H + zh) * W + zw;
                                float v = xp[ii];
                                if (bi < 0 || v > best || v != v) { best = v; bi = ii; }
                            }
                    int64_t o = nc * Do * Ho * Wo + ((int64_t)d * Ho + h) * Wo + w0;
                    y[o] = best;
                    if (idx) idx[o] = bi;
                }
    }
}

void e3o_maxpool_bwd(const float *dy, const int64_t *idx, float *dx,
                     int N, int C, int D, int H, int W, int kd, int kh, int kw)
{
    int Do = (D + kd - 1) / kd, Ho = (H + kh - 1) / kh, Wo = (W + kw - 1) / kw;
    memset(dx, 0, sizeof(float) * (size_t)N * C * D * H * W);
    for (int64_t nc = 0; nc < (int64_t)N * C; nc++)
        for (int64_t o = 0; o < (int64_t)Do * Ho * Wo; o++)
            dx[nc * D * H * W + idx[nc * Do * Ho * Wo + o]] += dy[nc * Do * Ho * Wo + o];
}
