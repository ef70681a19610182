/* TEST INFRASTRUCTURE ONLY -- NOT PRODUCT CODE.  See pqa_oracle.h for the scope statement and the parity pin.
 *
 * Scalar restatement of the reference's AVX2 CpuEngine<SRDoubleNumber> hot path. Every AVX lane operation of
 * the reference is one scalar IEEE operation here; build with -ffp-contract=off so that no multiply-add is
 * fused except where the reference itself uses _mm256_fmadd_pd (Log2Hot), which is written as fma().
 * Citations are relative to /root/reference/ProbQA/. */
#include "pqa_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------ helpers */
static inline uint64_t u64_of(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double f64_of(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

static inline int bit_test(const uint8_t *bits, int64_t i) {
  return bits ? ((bits[i >> 3] >> (i & 7)) & 1) : 0;
}
/* gap mask of lane j: removed target, or a padding lane of the last vector (GapTracker.h:12-15) */
static inline int lane_gap(const uint8_t *tgaps, int64_t j, int64_t T) {
  return (j >= T) ? 1 : bit_test(tgaps, j);
}

/* scalar Kahan: SRPlatform/Interface/SRAccumulator.h:15-41 */
typedef struct { double sum, corr; } K1;
static inline void k1_init(K1 *a, double v) { a->sum = v; a->corr = 0; }
static inline void k1_add(K1 *a, double v) {
  const double y = v - a->corr;
  const double t = a->sum + y;
  a->corr = (t - a->sum) - y;
  a->sum = t;
}
static inline void k1_neg(K1 *a) { a->sum = -a->sum; a->corr = -a->corr; }
static inline double k1_get(const K1 *a) { return a->sum - a->corr; }

/* ------------------------------------------------------------------ V4 accumulator */
void ora_v4_reset(OraV4 *a) { memset(a, 0, sizeof(*a)); }

void ora_v4_add(OraV4 *a, const double v[4]) {           /* SRAccumVectDbl256.h:40-46 */
  for (int c = 0; c < 4; c++) {
    const double y = v[c] - a->corr[c];
    const double t = a->sum[c] + y;
    a->corr[c] = (t - a->sum[c]) - y;
    a->sum[c] = t;
  }
}

void ora_v4_add_at(OraV4 *a, int at, double v) {          /* SRAccumVectDbl256.h:48-54 */
  const double y = v - a->corr[at];
  const double t = a->sum[at] + y;
  a->corr[at] = (t - a->sum[at]) - y;
  a->sum[at] = t;
}

double ora_v4_precise_sum(const OraV4 *a) {               /* SRAccumVectDbl256.h:83-91 */
  K1 ans; k1_init(&ans, a->corr[3]);
  for (int i = 2; i >= 0; i--) k1_add(&ans, a->corr[i]);
  k1_neg(&ans);
  for (int i = 3; i >= 0; i--) k1_add(&ans, a->sum[i]);
  return k1_get(&ans);
}

double ora_v4_pair_sum(const OraV4 *a, const OraV4 *f, double *fellowSum) { /* SRAccumVectDbl256.h:115-132 */
  /* Two independent SSE lanes doing what precise_sum does; lane 0 = this, lane 1 = fellow. */
  *fellowSum = ora_v4_precise_sum(f);
  return ora_v4_precise_sum(a);
}

double ora_v4_full_sum(const OraV4 *a) {                   /* SRAccumVectDbl256.h:56-60 */
  /* hadd(corr,sum) = [c0+c1, s0+s1, c2+c3, s2+s3]; upper128 + lower128; [1]-[0] */
  const double cs = (a->corr[2] + a->corr[3]) + (a->corr[0] + a->corr[1]);
  const double ss = (a->sum[2] + a->sum[3]) + (a->sum[0] + a->sum[1]);
  return ss - cs;
}

/* ------------------------------------------------------------------ Log2Hot */
static double g_log2tbl[1024];
static pthread_once_t g_tbl_once = PTHREAD_ONCE_INIT;

static void init_log2tbl(void) {                          /* SRPlatform/SRVectMath.cpp:30-44 */
  const double cCorr1 = 9.9999999999999927e-01;
  for (uint32_t i = 0; i < 1024; i++) {
    const uint64_t iZ = 0x3FF0000000000000ULL | ((uint64_t)i << (52 - 10));
    const uint64_t iZp = iZ | (1ULL << (52 - 10 - 1));
    g_log2tbl[i] = log2(f64_of(iZp));
  }
  g_log2tbl[0] *= cCorr1;
}

const double *ora_log2hot_table(void) { pthread_once(&g_tbl_once, init_log2tbl); return g_log2tbl; }

double ora_log2hot(double x) {                            /* SRPlatform/Interface/SRVectMath.h:87-135 */
  const double *tbl = ora_log2hot_table();
  const uint64_t bits = u64_of(x);
  /* z: exponent field replaced by 1023; the sign bit is kept (SRVectMath.cpp:17-18: mask clears only the exponent) */
  const double z = f64_of((bits & ~0x7FF0000000000000ULL) | 0x3FF0000000000000ULL);
  const int32_t high32 = (int32_t)(bits >> 32);
  const int32_t exps32 = high32 >> 20;                    /* arithmetic shift, sign not cleared (:96-98) */
  const int32_t normExp = exps32 - 1023;
  const uint32_t idx = (uint32_t)(high32 >> 10) & 1023u;  /* :101-102 */
  const double y = tbl[idx];
  /* exp2_Y = plusBit | (z & ~(2^42-1))  (:108, constants SRVectMath.cpp:19-22) */
  const double e2y = f64_of((1ULL << 41) | (u64_of(z) & ~((1ULL << 42) - 1)));
  const double tNum = z - e2y;
  const double tDen = z + e2y;
  const double t = tNum / tDen;
  const double t2 = t * t;
  const double t3 = t * t2;
  const double terms01 = fma(1.0 / 3, t3, t);
  const double log2_z = fma(terms01, 2.8853900817779268147198493620038, y);
  return log2_z + (double)normExp;
}

/* ------------------------------------------------------------------ split */
int64_t ora_calc_split(int64_t nItems, int64_t nWorkers, int64_t *bounds) { /* SRPoolRunner.h:96-110 */
  int64_t n = 0, next = 0;
  const int64_t quot = nItems / nWorkers, rem = nItems % nWorkers;
  while (n < nWorkers && next < nItems) {
    next += quot + ((n < rem) ? 1 : 0);
    bounds[n++] = next;
  }
  return n;
}

/* Sum of per-piece PreciseSum()s with a scalar Kahan (Summator.h:11-21), pieces over ceil(T/4) vectors. */
static double piecewise_sum(const double *m, int64_t T, int64_t W) {
  const int64_t nVects = (T + 3) >> 2;
  int64_t *bounds = (int64_t *)malloc(sizeof(int64_t) * (size_t)(W > 0 ? W : 1));
  const int64_t nPieces = ora_calc_split(nVects, W, bounds);
  K1 acc; k1_init(&acc, 0);
  int64_t first = 0;
  for (int64_t p = 0; p < nPieces; p++) {
    OraV4 a; ora_v4_reset(&a);
    for (int64_t v = first; v < bounds[p]; v++) {
      double lanes[4];
      for (int c = 0; c < 4; c++) { const int64_t j = 4 * v + c; lanes[c] = (j < T) ? m[j] : 0.0; }
      ora_v4_add(&a, lanes);
    }
    k1_add(&acc, ora_v4_precise_sum(&a));
    first = bounds[p];
  }
  free(bounds);
  return k1_get(&acc);
}

/* ------------------------------------------------------------------ StartQuiz */
void ora_start_quiz(const double *vB, const uint8_t *tgaps, int64_t T, int64_t W, double *prior) {
  for (int64_t j = 0; j < T; j++) prior[j] = bit_test(tgaps, j) ? 0.0 : vB[j];   /* CESetPriorsSubtaskSum.cpp:30-35 */
  const double S = piecewise_sum(prior, T, W);
  for (int64_t j = 0; j < T; j++) prior[j] = prior[j] / S;                         /* CEDivTargPriorsSubtask.h:16-21 */
}

/* ------------------------------------------------------------------ RecordAnswer */
void ora_record_answer(const double *sArow, const double *mDrow, const uint8_t *tgaps, int64_t T, int64_t W,
                       double *prior) {
  for (int64_t j = 0; j < T; j++) {                        /* CERecordAnswerSubtaskMul.cpp:27-37 */
    const double P = sArow[j] / mDrow[j];
    const double product = prior[j] * P;
    prior[j] = bit_test(tgaps, j) ? 0.0 : product;
  }
  const double S = piecewise_sum(prior, T, W);
  for (int64_t j = 0; j < T; j++) prior[j] = prior[j] / S;
}

/* ------------------------------------------------------------------ ResumeQuiz */
static uint8_t ceil_log2_u64(uint64_t v);

int ora_resume_quiz(const double *sA, const double *mD, const double *vB, int64_t ldT, int64_t K, int64_t T, int64_t W,
                    const uint8_t *tgaps, const OraAnsweredQuestion *aqs, int64_t nAQs, double *prior) {
  const uint64_t EXPMASK = 0x7FF0000000000000ULL, EXP0 = 0x3FF0000000000000ULL;
  const int64_t Tp = (T + 3) & ~(int64_t)3;
  double *mant = (double *)malloc(sizeof(double) * (size_t)Tp);
  int64_t *exps = (int64_t *)malloc(sizeof(int64_t) * (size_t)Tp);
  for (int64_t j = 0; j < Tp; j++) {                       /* CEUpdatePriorsSubtaskMul.cpp:40-82 */
    double m = 0; int64_t e = 0;
    for (int64_t i = 0; i < nAQs; i++) {
      const int64_t q = aqs[i].iQuestion, a = aqs[i].iAnswer;
      /* padding lanes of the last vector: the reference's rows are padded (A pad / D pad); their values never reach the
       * result because padding lanes are gaps in the normalisation below */
      const double A = (j < T) ? sA[(q * K + a) * ldT + j] : 0.0, D = (j < T) ? mD[q * ldT + j] : 1.0;
      const double P = A / D;                              /* :46, :65 */
      const double old = (i == 0) ? vB[j & 3] : m;         /* :48 loads pvB (vector 0) for every j; :68 */
      const double product = old * P;                      /* :49, :69 */
      const uint64_t pb = u64_of(product);
      m = f64_of((pb & ~EXPMASK) | EXP0);                  /* MakeExponent0, SRSimd.h:194-198 */
      const int64_t pe = (int64_t)((pb & EXPMASK) >> 52);  /* ExtractExponents64<false>, SRSimd.h:130-135 */
      e = (i == 0) ? pe : e + pe;                          /* :54-55, :74-77 */
    }
    mant[j] = m; exps[j] = e;
  }
  int64_t fullMax = INT64_MIN;                             /* CENormPriorsSubtaskMax.cpp:21-29,43-49 */
  for (int64_t j = 0; j < Tp; j++) {
    if (lane_gap(tgaps, j, T)) continue;
    const int64_t tot = exps[j] + (int64_t)((u64_of(mant[j]) & EXPMASK) >> 52);
    if (tot > fullMax) fullMax = tot;
  }
  const int64_t highBound = 1023 + 1023 - (int64_t)ceil_log2_u64((uint64_t)T) - 2;   /* CpuEngine.cpp:314 */
  const int64_t minAllowed = INT64_MIN + highBound + 1;
  if (fullMax <= minAllowed) { free(mant); free(exps); return 1; }                    /* :316-319 */
  const int64_t corr = highBound - fullMax;                /* :320 */
  for (int64_t j = 0; j < Tp; j++) {                       /* CENormPriorsSubtaskCorrSum.cpp:24-41 */
    const uint64_t mb = u64_of(mant[j]);
    const int64_t normExp = exps[j] + (int64_t)((mb & EXPMASK) >> 52) + corr;
    const int zero = (normExp < 1) || lane_gap(tgaps, j, T);
    const double nm = zero ? 0.0 : f64_of((mb & ~EXPMASK) | ((uint64_t)normExp << 52));   /* ReplaceExponents, SRSimd.h:200-204 */
    if (j < T) prior[j] = nm;
  }
  const double S = piecewise_sum(prior, T, W);             /* :57-62 + Summator.h:11-21 */
  for (int64_t j = 0; j < T; j++) prior[j] = prior[j] / S; /* CEDivTargPriorsSubtask.h:16-21 */
  free(mant); free(exps);
  return 0;
}

/* ------------------------------------------------------------------ question evaluation */
static double calc_velocity_component(double V, int64_t nTargets) { /* CEEvalQsSubtaskConsider.cpp:24-34 */
  const double cLnMaxV = 0.34657359027997265470861606072909;       /* SRMath::_cLnSqrt2 */
  const double lnV = (V == 0) ? -746.0 : log(V);
  const double powT = (double)nTargets * nTargets;
  return 1 / (cLnMaxV - lnV + cLnMaxV / powT);
}

double ora_eval_question(const double *sAi, const double *mDi, int64_t ldT, const double *prior,
                         const uint8_t *tgaps, int64_t K, int64_t T, int64_t nValidTargets,
                         double *Wk_out, double *Hk_out, double *Vk_out, double *lack_out, double *totW_out) {
  const int64_t nTV = (T + 3) >> 2, Tp = nTV * 4;
  double *invD = (double *)malloc(sizeof(double) * (size_t)Tp * 2);
  double *post = invD + Tp;
  double *Wk = (double *)malloc(sizeof(double) * (size_t)K * 3);
  double *Hk = Wk + K, *Vk = Hk + K;

  K1 accTotW; k1_init(&accTotW, 0);
  OraV4 accL; ora_v4_reset(&accL);
  for (int64_t k = 0; k < K; k++) {
    const double *sAik = sAi + k * ldT;
    OraV4 accLhEnt; ora_v4_reset(&accLhEnt);
    for (int64_t v = 0; v < nTV; v++) {                   /* pass 1: CEEvalQsSubtaskConsider.cpp:66-87 */
      double lik[4];
      for (int c = 0; c < 4; c++) {
        const int64_t j = 4 * v + c;
        const int gap = lane_gap(tgaps, j, T);
        if (k == 0) invD[j] = gap ? 0.0 : 1.0 / mDi[j];   /* :72-76 */
        const double P = gap ? 0.0 : sAik[j] * invD[j];   /* :81 (gap lanes are masked to +0 right after) */
        lik[c] = gap ? 0.0 : P * prior[j];                /* :82 */
        post[j] = lik[c];
      }
      ora_v4_add(&accLhEnt, lik);                         /* :86 */
    }
    const double W = ora_v4_precise_sum(&accLhEnt);       /* :88 */
    k1_add(&accTotW, W);
    Wk[k] = W;
    const double invWk = 1.0 / W;                         /* :91 */

    ora_v4_reset(&accLhEnt);
    OraV4 accV; ora_v4_reset(&accV);
    for (int64_t v = 0; v < nTV; v++) {                   /* pass 2: :95-128 */
      double H[4], L[4], V[4];
      for (int c = 0; c < 4; c++) {
        const int64_t j = 4 * v + c;
        const int gap = lane_gap(tgaps, j, T);
        const double posterior = post[j] * invWk;         /* :97 */
        const double pr = gap ? 0.0 : prior[j];           /* :103 */
        const double l2 = gap ? 0.0 : ora_log2hot(posterior); /* :106 */
        H[c] = posterior * l2;                            /* :113 */
        L[c] = gap ? 0.0 : (invD[j] * invD[j]) / l2;      /* :117 */
        const double diff = posterior - pr;               /* :119 */
        V[c] = diff * diff;                               /* :126 */
      }
      ora_v4_add(&accLhEnt, H);
      ora_v4_add(&accL, L);
      ora_v4_add(&accV, V);
    }
    double velocity;
    Hk[k] = -ora_v4_pair_sum(&accLhEnt, &accV, &velocity); /* :129-132 */
    Vk[k] = velocity;
  }
  const double totW = k1_get(&accTotW);                   /* :134 */

  OraV4 accAvgH, accAvgV; ora_v4_reset(&accAvgH); ora_v4_reset(&accAvgV);
  const int64_t nVectorized = (K >> 2) << 2;              /* :141-142 */
  for (int64_t k = 0; k < nVectorized; k += 4) {          /* :148-159 */
    double wH[4], wV[4];
    for (int c = 0; c < 4; c++) { wH[c] = Wk[k + c] * Hk[k + c]; wV[c] = Wk[k + c] * sqrt(Vk[k + c]); }
    ora_v4_add(&accAvgH, wH);
    ora_v4_add(&accAvgV, wV);
  }
  for (int64_t k = nVectorized; k < K; k++) {             /* :163-172 */
    const double velocity = sqrt(Vk[k]);
    ora_v4_add_at(&accAvgH, (int)(k - nVectorized), Wk[k] * Hk[k]);
    ora_v4_add_at(&accAvgV, (int)(k - nVectorized), Wk[k] * velocity);
  }
  double sumV;
  const double sumH = ora_v4_pair_sum(&accAvgH, &accAvgV, &sumV); /* :175 */
  const double avgH = sumH / totW;                        /* :176-177 */
  const double avgV = sumV / totW;
  const double nExpectedTargets = exp2(avgH);             /* :181 */
  const double vComp = calc_velocity_component(avgV, nValidTargets + 1); /* :191 */
  const double lack = -ora_v4_precise_sum(&accL);         /* :201 */
  const double priority = pow(lack, 1) * pow(vComp, 9) * pow(nExpectedTargets, -2); /* :207 */

  if (Wk_out) memcpy(Wk_out, Wk, sizeof(double) * (size_t)K);
  if (Hk_out) memcpy(Hk_out, Hk, sizeof(double) * (size_t)K);
  if (Vk_out) memcpy(Vk_out, Vk, sizeof(double) * (size_t)K);
  if (lack_out) *lack_out = lack;
  if (totW_out) *totW_out = totW;
  free(Wk); free(invD);
  return priority;
}

typedef struct {
  const double *sA, *mD, *prior; int64_t ldT;
  const uint8_t *asked, *qgaps, *tgaps;
  int64_t Q, K, T, nValid;
  double *runLength, *priority;
  const int64_t *bounds; int64_t nChunks;
  int64_t chunkFirst, chunkLimit; /* chunk indices handled by this thread */
} EvalCtx;

static void eval_chunk(const EvalCtx *c, int64_t ch) {     /* CEEvalQsSubtaskConsider.cpp:52-58,212-214 */
  const int64_t iFirst = ch ? c->bounds[ch - 1] : 0, iLimit = c->bounds[ch];
  K1 run; k1_init(&run, 0);
  for (int64_t i = iFirst; i < iLimit; i++) {
    if (bit_test(c->qgaps, i) || bit_test(c->asked, i)) {
      c->runLength[i] = k1_get(&run);
      if (c->priority) c->priority[i] = NAN;
      continue;
    }
    const double pr = ora_eval_question(c->sA + (size_t)i * c->K * c->ldT, c->mD + (size_t)i * c->ldT, c->ldT,
                                        c->prior, c->tgaps, c->K, c->T, c->nValid, NULL, NULL, NULL, NULL, NULL);
    if (c->priority) c->priority[i] = pr;
    k1_add(&run, pr);
    c->runLength[i] = k1_get(&run);
  }
}

static void *eval_thread(void *p) {
  const EvalCtx *c = (const EvalCtx *)p;
  for (int64_t ch = c->chunkFirst; ch < c->chunkLimit; ch++) eval_chunk(c, ch);
  return NULL;
}

int64_t ora_eval_questions(const double *sA, const double *mD, int64_t ldT, const double *prior,
                           const uint8_t *asked, const uint8_t *qgaps, const uint8_t *tgaps,
                           int64_t Q, int64_t K, int64_t T, int64_t W, int nThreads,
                           double *runLength, double *priority, double *grand, int64_t *bounds) {
  ora_log2hot_table();
  const int64_t nWorkers = W * 8;                          /* CpuEngine.cpp:339 */
  const int64_t nChunks = ora_calc_split(Q, nWorkers, bounds);
  int64_t nGaps = 0;
  if (tgaps) for (int64_t j = 0; j < T; j++) nGaps += bit_test(tgaps, j);
  EvalCtx base = { sA, mD, prior, ldT, asked, qgaps, tgaps, Q, K, T, T - nGaps /* CpuEngine.cpp:351 */,
                   runLength, priority, bounds, nChunks, 0, nChunks };
  if (nThreads <= 1) {
    eval_thread(&base);
  } else {
    if (nThreads > nChunks) nThreads = (int)nChunks;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nThreads);
    EvalCtx *ctx = (EvalCtx *)malloc(sizeof(EvalCtx) * (size_t)nThreads);
    for (int t = 0; t < nThreads; t++) {
      ctx[t] = base;
      ctx[t].chunkFirst = nChunks * t / nThreads;
      ctx[t].chunkLimit = nChunks * (t + 1) / nThreads;
      pthread_create(&th[t], NULL, eval_thread, &ctx[t]);
    }
    for (int t = 0; t < nThreads; t++) pthread_join(th[t], NULL);
    free(ctx); free(th);
  }
  K1 accTotG; k1_init(&accTotG, 0);                        /* CpuEngine.cpp:362-374 */
  for (int64_t c = 0; c < nChunks; c++) {
    k1_add(&accTotG, runLength[bounds[c] - 1]);
    grand[c] = k1_get(&accTotG);
  }
  return nChunks;
}

/* ------------------------------------------------------------------ selection */
static uint64_t pack64(const uint8_t *bits, int64_t iPack, int64_t nBits, int padOnes) {
  /* 64 bits starting at bit 64*iPack; bits >= nBits read as padOnes (gaps: 1, asked: 0) */
  uint64_t r = 0;
  for (int b = 0; b < 64; b++) {
    const int64_t i = iPack * 64 + b;
    int v;
    if (i >= nBits) v = padOnes; else v = bit_test(bits, i);
    r |= (uint64_t)v << b;
  }
  return r;
}

int64_t ora_find_nearest_question(int64_t iMiddle, int64_t Q, const uint8_t *asked, const uint8_t *qgaps) {
  /* BaseEngine.cpp:60-124 */
  const uint32_t dInf = 200;
  const int64_t iPack64 = iMiddle >> 6;
  const uint32_t iWithin = (uint32_t)(iMiddle & 63);
#define AVAIL(p) (~(pack64(qgaps, (p), Q, 1) | pack64(asked, (p), Q, 0)))
  const uint64_t available = AVAIL(iPack64);
  if (available != 0) {
    const uint64_t baseMask = (1ULL << iWithin) - 1;
    const uint64_t higher = available & ~baseMask;
    const uint64_t lower = baseMask & available;
    const uint32_t dHigher = higher ? ((uint32_t)__builtin_ctzll(higher) - iWithin) : dInf;
    const uint32_t dLower = lower ? (iWithin - (uint32_t)(63 - __builtin_clzll(lower))) : dInf;
    if (dHigher < dLower) return iMiddle + dHigher;
    return iMiddle - dLower;
  }
  const int64_t limPack64 = (Q + 63) >> 6;
  int64_t i = 1;
  while ((iPack64 >= i) && (iPack64 + i < limPack64)) {
    const uint64_t availLeft = AVAIL(iPack64 - i);
    const uint64_t availRight = AVAIL(iPack64 + i);
    if ((availLeft | availRight) == 0) { i++; continue; }
    const uint32_t dHigher = availRight ? ((uint32_t)__builtin_ctzll(availRight) + 64 - iWithin) : dInf;
    const uint32_t dLower = availLeft ? (iWithin + 64 - (uint32_t)(63 - __builtin_clzll(availLeft))) : dInf;
    if (dHigher < dLower) return iMiddle + dHigher + ((i - 1) << 6);
    return iMiddle - dLower - ((i - 1) << 6);
  }
  while (iPack64 >= i) {
    const uint64_t availLeft = AVAIL(iPack64 - i);
    if (!availLeft) { i++; continue; }
    const uint32_t dLower = iWithin + 64 - (uint32_t)(63 - __builtin_clzll(availLeft));
    return iMiddle - dLower - ((i - 1) << 6);
  }
  while (iPack64 + i < limPack64) {
    const uint64_t availRight = AVAIL(iPack64 + i);
    if (!availRight) { i++; continue; }
    const uint32_t dHigher = (uint32_t)__builtin_ctzll(availRight) + 64 - iWithin;
    return iMiddle + dHigher + ((i - 1) << 6);
  }
#undef AVAIL
  return -1;
}

static int64_t upper_bound_d(const double *a, int64_t n, double v) { /* std::upper_bound: first a[i] > v */
  int64_t lo = 0, len = n;
  while (len > 0) {
    const int64_t half = len >> 1;
    if (!(v < a[lo + half])) { lo += half + 1; len -= half + 1; } else len = half;
  }
  return lo;
}

int64_t ora_select_question(const double *runLength, const double *grand, const int64_t *bounds, int64_t nChunks,
                            int64_t Q, uint64_t rnd, const uint8_t *asked, const uint8_t *qgaps) {
  const double totG = grand[nChunks - 1];                  /* CpuEngine.cpp:375 */
  /* SRDoubleNumber.h:35-39: upper * u64 / numeric_limits<uint64_t>::max() (left-assoc, both converted to double) */
  const double selRunLen = totG * (double)rnd / (double)UINT64_MAX;
  int64_t sel;
  const int64_t iWorker = upper_bound_d(grand, nChunks, selRunLen); /* :380-381 */
  if (iWorker >= nChunks) {
    sel = Q - 1;                                           /* :382-386 */
  } else {
    const double inWorker = selRunLen - (iWorker == 0 ? 0.0 : grand[iWorker - 1]); /* :388 */
    const int64_t iFirst = iWorker == 0 ? 0 : bounds[iWorker - 1], iLimit = bounds[iWorker];
    sel = iFirst + upper_bound_d(runLength + iFirst, iLimit - iFirst, inWorker);   /* :391 */
    if (sel >= iLimit) sel = iLimit - 1;                   /* :392-400 */
  }
  if (bit_test(qgaps, sel) || bit_test(asked, sel)) sel = ora_find_nearest_question(sel, Q, asked, qgaps); /* :404-406 */
  return sel;
}

/* ------------------------------------------------------------------ heaps (libstdc++ bits/stl_heap.h, max-heap on prob) */
#define HLESS(a, b) ((a).prob < (b).prob)                  /* RatedTarget::operator<  Interface/PqaCommon.h:58-60 */

static void rt_push_heap(OraRatedTarget *first, int64_t hole, int64_t top, OraRatedTarget value) {
  int64_t parent = (hole - 1) / 2;
  while (hole > top && HLESS(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

static void rt_adjust_heap(OraRatedTarget *first, int64_t hole, int64_t len, OraRatedTarget value) {
  const int64_t top = hole;
  int64_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (HLESS(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  rt_push_heap(first, hole, top, value);
}

void ora_make_heap(OraRatedTarget *first, int64_t len) {
  if (len < 2) return;
  int64_t parent = (len - 2) / 2;
  for (;;) {
    OraRatedTarget value = first[parent];
    rt_adjust_heap(first, parent, len, value);
    if (parent == 0) return;
    parent--;
  }
}

void ora_pop_heap(OraRatedTarget *first, int64_t len) {
  if (len > 1) {
    OraRatedTarget value = first[len - 1];
    first[len - 1] = first[0];
    rt_adjust_heap(first, 0, len - 1, value);
  }
}

/* head heap item: RatingsHeap.h:11-20 (prob first, then source) -- same algorithms, different payload */
typedef struct { double prob; int64_t iSource; } HeadItem;
static void hh_push_heap(HeadItem *first, int64_t hole, int64_t top, HeadItem value) {
  int64_t parent = (hole - 1) / 2;
  while (hole > top && HLESS(first[parent], value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
  first[hole] = value;
}
static void hh_adjust_heap(HeadItem *first, int64_t hole, int64_t len, HeadItem value) {
  const int64_t top = hole;
  int64_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (HLESS(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); first[hole] = first[child - 1]; hole = child - 1; }
  hh_push_heap(first, hole, top, value);
}
static void hh_make_heap(HeadItem *first, int64_t len) {
  if (len < 2) return;
  for (int64_t parent = (len - 2) / 2;; parent--) {
    HeadItem value = first[parent];
    hh_adjust_heap(first, parent, len, value);
    if (parent == 0) return;
  }
}
static void hh_pop_heap(HeadItem *first, int64_t len) {
  if (len > 1) { HeadItem value = first[len - 1]; first[len - 1] = first[0]; hh_adjust_heap(first, 0, len - 1, value); }
}
static void hh_down(HeadItem *first, int64_t len) {       /* SRPlatform/Interface/SRHeap.h:16-37 */
  int64_t cur = 0;
  for (;;) {
    const int64_t c1 = 2 * cur + 1;
    if (c1 >= len) return;
    const int64_t c2 = c1 + 1;
    if (c2 >= len) {
      if (HLESS(first[cur], first[c1])) { HeadItem t = first[cur]; first[cur] = first[c1]; first[c1] = t; }
      return;
    }
    const int64_t hi = HLESS(first[c2], first[c1]) ? c1 : c2;
    if (!HLESS(first[cur], first[hi])) return;
    HeadItem t = first[cur]; first[cur] = first[hi]; first[hi] = t;
    cur = hi;
  }
}

static uint8_t ceil_log2_u64(uint64_t v) {                 /* SRMath::CeilLog2 */
  if (v <= 1) return 0;
  return (uint8_t)(64 - __builtin_clzll(v - 1));
}

int ora_would_use_radix(int64_t T, int64_t W, int64_t maxCount) { /* CpuEngine.cpp:423-432 */
  const uint64_t nTargPerThread = ((uint64_t)T + (uint64_t)W - 1) / (uint64_t)W;
  uint64_t a = nTargPerThread > 256 ? nTargPerThread : 256;
  uint8_t lw = ceil_log2_u64((uint64_t)W); if (lw < 1) lw = 1;
  const uint64_t nRadix = 9 * a + (uint64_t)maxCount * lw;
  const uint64_t nHeapify = 3 * nTargPerThread + (uint64_t)maxCount * ceil_log2_u64((uint64_t)T);
  return nRadix < nHeapify;
}

int64_t ora_list_top_targets(const double *prior, const uint8_t *tgaps, int64_t T, int64_t W, int64_t maxCount,
                             OraRatedTarget *dest) {
  int64_t *bounds = (int64_t *)malloc(sizeof(int64_t) * (size_t)W * 2);
  int64_t *limits = bounds + W;
  OraRatedTarget *ratings = (OraRatedTarget *)malloc(sizeof(OraRatedTarget) * (size_t)T);
  HeadItem *head = (HeadItem *)malloc(sizeof(HeadItem) * (size_t)W);
  const int64_t nPieces = ora_calc_split(T, W, bounds);   /* CEListTopTargetsAlgorithm.cpp:44 */
  int64_t first = 0;
  for (int64_t p = 0; p < nPieces; p++) {                  /* CEHeapifyPriorsSubtaskMake.cpp:42-53,56-87 */
    int64_t sel = first;
    for (int64_t j = first; j < bounds[p]; j++) {
      if (bit_test(tgaps, j)) continue;
      const double prob = prior[j];
      if (prob <= 0) continue;
      ratings[sel].prob = prob; ratings[sel].iTarget = j; sel++;
    }
    limits[p] = sel;
    ora_make_heap(ratings + first, sel - first);
    first = bounds[p];
  }
  for (int64_t p = nPieces - 1; p >= 1; p--) bounds[p] = bounds[p - 1]; /* RecalcToStarts SRPoolRunner.h:70-76 */
  if (nPieces > 0) bounds[0] = 0;
  int64_t nHh = 0;
  for (int64_t p = 0; p < nPieces; p++) {                  /* CEListTopTargetsAlgorithm.cpp:57-66 */
    if (limits[p] == bounds[p]) continue;
    head[nHh].iSource = p; head[nHh].prob = ratings[bounds[p]].prob; nHh++;
  }
  hh_make_heap(head, nHh);
  int64_t listed = maxCount;
  for (int64_t i = 0; i < maxCount; i++) {                 /* :68-94 */
    if (nHh == 0) { listed = i; break; }
    dest[i].prob = head[0].prob;
    const int64_t piece = head[0].iSource;
    const int64_t start = bounds[piece];
    dest[i].iTarget = ratings[start].iTarget;
    const int64_t lim = limits[piece];
    if (start + 1 == lim) { hh_pop_heap(head, nHh); nHh--; continue; }
    ora_pop_heap(ratings + start, lim - start);
    limits[piece]--;
    head[0].prob = ratings[start].prob;
    hh_down(head, nHh);
  }
  free(head); free(ratings); free(bounds);
  return listed;
}

/* ------------------------------------------------------------------ training */
typedef struct { double *sA, *mD; int64_t ldT, K, iTarget; double inc2B, incBSquare, inc4B, incSquare2B; } TrainOp;
#define A_AT(op, q, a) ((op)->sA[((size_t)(q) * (op)->K + (a)) * (op)->ldT + (op)->iTarget])
#define D_AT(op, q) ((op)->mD[(size_t)(q) * (op)->ldT + (op)->iTarget])

static void train_process_one(TrainOp *op, const OraAnsweredQuestion *aq, double twoB, double bSquare) {
  /* CETrainOperation.cpp:15-26 */
  const double aSquare = A_AT(op, aq->iQuestion, aq->iAnswer);
  const double a = sqrt(aSquare);
  const double addend = a * twoB + bSquare;
  A_AT(op, aq->iQuestion, aq->iAnswer) = aSquare + addend;
  D_AT(op, aq->iQuestion) = D_AT(op, aq->iQuestion) + addend;
}

static void train_perform2(TrainOp *op, const OraAnsweredQuestion *f, const OraAnsweredQuestion *s) {
  /* CETrainOperation.cpp:32-83 */
  if (f->iQuestion == s->iQuestion) {
    if (f->iAnswer == s->iAnswer) {
      train_process_one(op, f, op->inc4B, op->incSquare2B);            /* :34-36 */
    } else {
      const double a0sq = A_AT(op, f->iQuestion, f->iAnswer), a1sq = A_AT(op, s->iQuestion, s->iAnswer);
      const double add0 = sqrt(a0sq) * op->inc2B + op->incBSquare;
      const double add1 = sqrt(a1sq) * op->inc2B + op->incBSquare;
      A_AT(op, f->iQuestion, f->iAnswer) = a0sq + add0;
      A_AT(op, s->iQuestion, s->iAnswer) = a1sq + add1;
      D_AT(op, f->iQuestion) = D_AT(op, f->iQuestion) + (add0 + add0);  /* :45-46 -- twice addend0, as the reference does */
    }
  } else {
    const double a0sq = A_AT(op, f->iQuestion, f->iAnswer), a1sq = A_AT(op, s->iQuestion, s->iAnswer);
    const double add0 = sqrt(a0sq) * op->inc2B + op->incBSquare;
    const double add1 = sqrt(a1sq) * op->inc2B + op->incBSquare;
    A_AT(op, f->iQuestion, f->iAnswer) = a0sq + add0;
    A_AT(op, s->iQuestion, s->iAnswer) = a1sq + add1;
    D_AT(op, f->iQuestion) = D_AT(op, f->iQuestion) + add0;
    D_AT(op, s->iQuestion) = D_AT(op, s->iQuestion) + add1;
  }
}

static void train_op_init(TrainOp *op, double *sA, double *mD, int64_t ldT, int64_t K, int64_t iTarget, double amount) {
  op->sA = sA; op->mD = mD; op->ldT = ldT; op->K = K; op->iTarget = iTarget;
  op->inc2B = 2 * amount; op->incBSquare = amount * amount;            /* CETrainTaskNumSpec.h:24-32 */
  op->inc4B = 4 * amount; op->incSquare2B = 4 * op->incBSquare;
}

void ora_record_quiz_target(double *sA, double *mD, double *vB, int64_t ldT, int64_t K,
                            const OraAnsweredQuestion *aqs, int64_t nAQs, int64_t iTarget, double amount) {
  TrainOp op; train_op_init(&op, sA, mD, ldT, K, iTarget, amount);
  int64_t i = 0;                                            /* CpuEngine.cpp:451-462 */
  const int64_t iEn = nAQs - 1;
  for (; i < iEn; i += 2) train_perform2(&op, &aqs[i], &aqs[i + 1]);
  if (i == iEn) train_process_one(&op, &aqs[i], op.inc2B, op.incBSquare);
  vB[iTarget] += amount;
}

void ora_train(double *sA, double *mD, double *vB, int64_t ldT, int64_t K,
               const OraAnsweredQuestion *aqs, int64_t nAQs, int64_t iTarget, double amount, int64_t W) {
  TrainOp op; train_op_init(&op, sA, mD, ldT, K, iTarget, amount);
  int64_t *last = (int64_t *)malloc(sizeof(int64_t) * (size_t)(W + (nAQs > 0 ? nAQs : 1)));
  int64_t *prev = last + W;
  for (int64_t w = 0; w < W; w++) last[w] = -1;
  for (int64_t i = 0; i < nAQs; i++) {                      /* CETrainSubtaskDistrib.h:45-51, tickets in order */
    const int64_t b = aqs[i].iQuestion % W;
    prev[i] = last[b]; last[b] = i;
  }
  for (int64_t w = 0; w < W; w++) {                         /* CETrainSubtaskAdd.cpp:17-38 */
    int64_t iLast = last[w];
    if (iLast == -1) continue;
    do {
      const OraAnsweredQuestion *f = &aqs[iLast];
      iLast = prev[iLast];
      if (iLast == -1) { train_process_one(&op, f, op.inc2B, op.incBSquare); break; }
      const OraAnsweredQuestion *s = &aqs[iLast];
      train_perform2(&op, f, s);
      iLast = prev[iLast];
    } while (iLast != -1);
  }
  free(last);
  vB[iTarget] += amount;                                    /* CpuEngine.cpp:175 */
}
