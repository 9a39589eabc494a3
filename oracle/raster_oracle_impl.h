/* Body of the C compositing oracle, included twice by raster_oracle.c (REAL = float, double).
 * TEST INFRASTRUCTURE ONLY -- see raster_oracle.c for what it restates. */

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(name, SUFFIX)

typedef struct {
  int C, N, D, W, H, tile_w, tile_h;
  const REAL *means2d, *conics, *colors, *opac, *ray_ts, *ray_planes, *normals, *Ks, *backgrounds;
  const int32_t* offsets;
  const int32_t* flatten_ids;
  int64_t M;
  /* forward outputs (also read by the backward) */
  REAL *out_colors, *out_alphas, *out_dexp, *out_dmed, *out_normals;
  int32_t *last_ids, *median_ids;
  uint8_t* fragile;
  int64_t* counters; /* {pairs visited, pairs blended} */
  /* backward inputs / outputs */
  const REAL *v_colors, *v_alphas, *v_dexp, *v_dmed, *v_normals;
  REAL *g_means2d, *g_conics, *g_colors, *g_opac, *g_ray_ts, *g_ray_planes, *g_normals, *g_backgrounds;
  int backward;
  double margin; /* scale of the 'fragile' margins (1 = the margins of rade_oracle.py) */
  int next_tile; /* work queue (atomic) */
} FN(Job);

static void FN(atomic_add)(REAL* p, REAL v) {
  if (v == (REAL)0) return;
#if REAL_IS_DOUBLE
  uint64_t* ip = (uint64_t*)p;
  uint64_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), neu;
  do {
    double f;
    memcpy(&f, &old, 8);
    f += v;
    memcpy(&neu, &f, 8);
  } while (!__atomic_compare_exchange_n(ip, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
#else
  uint32_t* ip = (uint32_t*)p;
  uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), neu;
  do {
    float f;
    memcpy(&f, &old, 4);
    f += v;
    memcpy(&neu, &f, 4);
  } while (!__atomic_compare_exchange_n(ip, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
#endif
}

/* One tile, forward: every pixel walks the tile's depth-sorted list front to back. */
static void FN(tile_fwd)(const FN(Job) * j, int tid, int64_t* n_tested, int64_t* n_contrib) {
  const int tiles = j->tile_w * j->tile_h;
  const int c = tid / tiles, tl = tid % tiles, ty = tl / j->tile_w, tx = tl % j->tile_w;
  const int64_t s = j->offsets[tid], e = (tid + 1 < j->C * tiles) ? j->offsets[tid + 1] : j->M;
  const int x0 = tx * RO_TILE, y0 = ty * RO_TILE;
  const int x1 = x0 + RO_TILE < j->W ? x0 + RO_TILE : j->W, y1 = y0 + RO_TILE < j->H ? y0 + RO_TILE : j->H;
  const int D = j->D;
  const REAL fx = j->Ks[c * 9 + 0], fy = j->Ks[c * 9 + 4], cx = j->Ks[c * 9 + 2], cy = j->Ks[c * 9 + 5];
  const REAL* bg = j->backgrounds ? j->backgrounds + (size_t)c * D : NULL;
  for (int yi = y0; yi < y1; ++yi)
    for (int xi = x0; xi < x1; ++xi) {
      const size_t pix = ((size_t)c * j->H + yi) * j->W + xi;
      const REAL px = (REAL)xi + (REAL)0.5, py = (REAL)yi + (REAL)0.5;
      const REAL rx = (px - cx) / fx, ry = (py - cy) / fy;
      const REAL ln = SQRT(rx * rx + ry * ry + (REAL)1);
      REAL T = 1, dsum = 0, tmed = 0, nrm[3] = {0, 0, 0};
      REAL* oc = j->out_colors + pix * D;
      for (int k = 0; k < D; ++k) oc[k] = 0;
      int32_t last = (int32_t)s - 1, med = -1;
      int frag = 0;
      for (int64_t i = s; i < e; ++i) {
        const int32_t g = j->flatten_ids[i];
        const REAL dx = j->means2d[2 * (size_t)g] - px, dy = j->means2d[2 * (size_t)g + 1] - py;
        const REAL* con = j->conics + 3 * (size_t)g;
        const REAL sigma = (REAL)0.5 * (con[0] * dx * dx + con[2] * dy * dy) + con[1] * dx * dy;
        const REAL a_raw = j->opac[g] * EXP(-sigma);
        const REAL alpha = a_raw < (REAL)RO_ALPHA_MAX ? a_raw : (REAL)RO_ALPHA_MAX;
        ++*n_tested;
        if (sigma >= 0 && FABS(alpha - (REAL)RO_ALPHA_MIN) < (REAL)(2e-4 * RO_ALPHA_MIN * j->margin)) frag = 1;
        if (!(sigma >= 0 && alpha >= (REAL)RO_ALPHA_MIN)) continue;
        const REAL nT = T * ((REAL)1 - alpha);
        if (FABS(nT - (REAL)RO_T_STOP) < (REAL)(2e-4 * RO_T_STOP * j->margin) || FABS(nT - (REAL)0.5) < (REAL)(1e-5 * j->margin)) frag = 1;
        if (!(nT > (REAL)RO_T_STOP)) break; /* the pixel is saturated: this Gaussian is not blended */
        const REAL vis = alpha * T;
        const REAL t = j->ray_ts[g] + j->ray_planes[2 * (size_t)g] * dx + j->ray_planes[2 * (size_t)g + 1] * dy;
        const REAL* col = j->colors + (size_t)g * D;
        for (int k = 0; k < D; ++k) oc[k] += vis * col[k];
        dsum += vis * t;
        for (int k = 0; k < 3; ++k) nrm[k] += vis * j->normals[3 * (size_t)g + k];
#if RO_MEDIAN_INCLUSIVE
        if (T > (REAL)0.5 && nT <= (REAL)0.5) { tmed = t; med = (int32_t)i; }
#else
        if (T > (REAL)0.5 && nT < (REAL)0.5) { tmed = t; med = (int32_t)i; }
#endif
        last = (int32_t)i;
        T = nT;
        ++*n_contrib;
      }
      if (bg)
        for (int k = 0; k < D; ++k) oc[k] += T * bg[k];
      j->out_alphas[pix] = (REAL)1 - T;
      j->out_dexp[pix] = dsum / ln;
      j->out_dmed[pix] = tmed / ln;
      for (int k = 0; k < 3; ++k) j->out_normals[pix * 3 + k] = nrm[k];
      j->last_ids[pix] = last;
      j->median_ids[pix] = med;
      if (j->fragile) j->fragile[pix] = (uint8_t)frag;
    }
}

/* One tile, backward (Appendix A9), written per pixel the way the chain rule reads: the forward is replayed to
 * record (alpha_i, T_i) of every contributing Gaussian, then the list is walked back to front with suffix sums
 * S_k = sum_{j > i} vis_j val_jk.  Per-Gaussian gradients are accumulated in a tile-local table first and added to
 * the global arrays once per (tile, Gaussian). */
static void FN(tile_bwd)(const FN(Job) * j, int tid, REAL** scratch, size_t* scratch_len) {
  const int tiles = j->tile_w * j->tile_h;
  const int c = tid / tiles, tl = tid % tiles, ty = tl / j->tile_w, tx = tl % j->tile_w;
  const int64_t s = j->offsets[tid], e = (tid + 1 < j->C * tiles) ? j->offsets[tid + 1] : j->M;
  if (e <= s) return;
  const int x0 = tx * RO_TILE, y0 = ty * RO_TILE;
  const int x1 = x0 + RO_TILE < j->W ? x0 + RO_TILE : j->W, y1 = y0 + RO_TILE < j->H ? y0 + RO_TILE : j->H;
  const int D = j->D;
  const int NG = 12 + D; /* xy 2, conic 3, opac 1, ray_t 1, ray_plane 2, normal 3, colours D */
  const size_t G = (size_t)(e - s);
  const size_t need = G * (size_t)NG + 2 * G;
  if (*scratch_len < need) {
    free(*scratch);
    *scratch = (REAL*)malloc(need * sizeof(REAL));
    *scratch_len = need;
  }
  REAL* acc = *scratch;            /* [G][NG] */
  REAL* h_alpha = acc + G * NG;    /* [G] alpha of the pairs of the current pixel (0 = not blended) */
  REAL* h_T = h_alpha + G;         /* [G] transmittance in front of them */
  memset(acc, 0, G * NG * sizeof(REAL));
  const REAL fx = j->Ks[c * 9 + 0], fy = j->Ks[c * 9 + 4], cx = j->Ks[c * 9 + 2], cy = j->Ks[c * 9 + 5];
  const REAL* bg = j->backgrounds ? j->backgrounds + (size_t)c * D : NULL;
  REAL* S = (REAL*)malloc((size_t)(D + 4) * sizeof(REAL));
  for (int yi = y0; yi < y1; ++yi)
    for (int xi = x0; xi < x1; ++xi) {
      const size_t pix = ((size_t)c * j->H + yi) * j->W + xi;
      const int64_t last = j->last_ids[pix];
      if (last < s) { /* nothing blended: only the background term */
        if (bg && j->g_backgrounds)
          for (int k = 0; k < D; ++k) FN(atomic_add)(j->g_backgrounds + (size_t)c * D + k, j->v_colors[pix * D + k]);
        continue;
      }
      const REAL px = (REAL)xi + (REAL)0.5, py = (REAL)yi + (REAL)0.5;
      const REAL rx = (px - cx) / fx, ry = (py - cy) / fy;
      const REAL iln = (REAL)1 / SQRT(rx * rx + ry * ry + (REAL)1);
      /* replay the forward up to the last blended Gaussian */
      REAL T = 1;
      for (int64_t i = s; i <= last; ++i) {
        const int32_t g = j->flatten_ids[i];
        const REAL dx = j->means2d[2 * (size_t)g] - px, dy = j->means2d[2 * (size_t)g + 1] - py;
        const REAL* con = j->conics + 3 * (size_t)g;
        const REAL sigma = (REAL)0.5 * (con[0] * dx * dx + con[2] * dy * dy) + con[1] * dx * dy;
        const REAL a_raw = j->opac[g] * EXP(-sigma);
        const REAL alpha = a_raw < (REAL)RO_ALPHA_MAX ? a_raw : (REAL)RO_ALPHA_MAX;
        if (!(sigma >= 0 && alpha >= (REAL)RO_ALPHA_MIN)) { h_alpha[i - s] = 0; continue; }
        h_alpha[i - s] = alpha;
        h_T[i - s] = T;
        T *= (REAL)1 - alpha;
      }
      const REAL T_final = T;
      const REAL* vc = j->v_colors + pix * D;
      const REAL v_al = j->v_alphas[pix];
      const REAL v_ds = j->v_dexp[pix] * iln;      /* d L / d (sum vis t) */
      const REAL v_dm = j->v_dmed[pix] * iln;      /* d L / d t_median */
      const REAL* vn = j->v_normals + pix * 3;
      const int64_t med = j->median_ids[pix];
      REAL bgdot = 0;
      if (bg)
        for (int k = 0; k < D; ++k) {
          bgdot += bg[k] * vc[k];
          if (j->g_backgrounds) FN(atomic_add)(j->g_backgrounds + (size_t)c * D + k, T_final * vc[k]);
        }
      /* d L / d T_final: alpha_out = 1 - T_final, colour += T_final * bg */
      const REAL v_Tfin = -v_al + bgdot;
      for (int k = 0; k < D + 4; ++k) S[k] = 0;
      for (int64_t i = last; i >= s; --i) {
        const REAL alpha = h_alpha[i - s];
        if (alpha == 0) continue;
        const REAL Ti = h_T[i - s];
        const int32_t g = j->flatten_ids[i];
        const REAL dx = j->means2d[2 * (size_t)g] - px, dy = j->means2d[2 * (size_t)g + 1] - py;
        const REAL* con = j->conics + 3 * (size_t)g;
        const REAL* rp = j->ray_planes + 2 * (size_t)g;
        const REAL* col = j->colors + (size_t)g * D;
        const REAL* nr = j->normals + 3 * (size_t)g;
        const REAL t = j->ray_ts[g] + rp[0] * dx + rp[1] * dy;
        const REAL vis = alpha * Ti;
        const REAL ra = (REAL)1 / ((REAL)1 - alpha);
        REAL* a = acc + (size_t)(i - s) * NG;
        /* values blended with weight vis: colours, t, normal */
        REAL v_alpha = 0;
        for (int k = 0; k < D; ++k) {
          v_alpha += vc[k] * (col[k] * Ti - S[k] * ra);
          a[12 + k] += vis * vc[k];
          S[k] += vis * col[k];
        }
        v_alpha += v_ds * (t * Ti - S[D] * ra);
        S[D] += vis * t;
        for (int k = 0; k < 3; ++k) {
          v_alpha += vn[k] * (nr[k] * Ti - S[D + 1 + k] * ra);
          a[9 + k] += vis * vn[k];
          S[D + 1 + k] += vis * nr[k];
        }
        /* T_final = prod (1 - alpha_j): d T_final / d alpha_i = -T_final / (1 - alpha_i) */
        v_alpha += v_Tfin * (-T_final * ra);
        REAL v_t = vis * v_ds;
        if (i == med) v_t += v_dm;
        a[6] += v_t;
        a[7] += v_t * dx;
        a[8] += v_t * dy;
        REAL gx = v_t * rp[0], gy = v_t * rp[1];
        /* alpha = min(o e^-sigma, ALPHA_MAX): no gradient through the clamp */
        const REAL sigma = (REAL)0.5 * (con[0] * dx * dx + con[2] * dy * dy) + con[1] * dx * dy;
        const REAL ex = EXP(-sigma);
        if (j->opac[g] * ex <= (REAL)RO_ALPHA_MAX) {
          const REAL v_sigma = -alpha * v_alpha;
          a[5] += ex * v_alpha;
          a[2] += (REAL)0.5 * v_sigma * dx * dx;
          a[3] += v_sigma * dx * dy;
          a[4] += (REAL)0.5 * v_sigma * dy * dy;
          gx += v_sigma * (con[0] * dx + con[1] * dy);
          gy += v_sigma * (con[1] * dx + con[2] * dy);
        }
        a[0] += gx;
        a[1] += gy;
      }
    }
  free(S);
  for (size_t i = 0; i < G; ++i) {
    const int32_t g = j->flatten_ids[s + (int64_t)i];
    const REAL* a = acc + i * NG;
    FN(atomic_add)(j->g_means2d + 2 * (size_t)g, a[0]);
    FN(atomic_add)(j->g_means2d + 2 * (size_t)g + 1, a[1]);
    for (int k = 0; k < 3; ++k) FN(atomic_add)(j->g_conics + 3 * (size_t)g + k, a[2 + k]);
    FN(atomic_add)(j->g_opac + g, a[5]);
    FN(atomic_add)(j->g_ray_ts + g, a[6]);
    FN(atomic_add)(j->g_ray_planes + 2 * (size_t)g, a[7]);
    FN(atomic_add)(j->g_ray_planes + 2 * (size_t)g + 1, a[8]);
    for (int k = 0; k < 3; ++k) FN(atomic_add)(j->g_normals + 3 * (size_t)g + k, a[9 + k]);
    for (int k = 0; k < D; ++k) FN(atomic_add)(j->g_colors + (size_t)g * D + k, a[12 + k]);
  }
}

static void* FN(worker)(void* arg) {
  FN(Job)* j = (FN(Job)*)arg;
  const int n_tiles = j->C * j->tile_w * j->tile_h;
  int64_t n_tested = 0, n_contrib = 0;
  REAL* scratch = NULL;
  size_t scratch_len = 0;
  for (;;) {
    const int tid = __atomic_fetch_add(&j->next_tile, 1, __ATOMIC_RELAXED);
    if (tid >= n_tiles) break;
    if (j->backward) FN(tile_bwd)(j, tid, &scratch, &scratch_len);
    else FN(tile_fwd)(j, tid, &n_tested, &n_contrib);
  }
  free(scratch);
  if (!j->backward && j->counters) {
    __atomic_fetch_add(&j->counters[0], n_tested, __ATOMIC_RELAXED);
    __atomic_fetch_add(&j->counters[1], n_contrib, __ATOMIC_RELAXED);
  }
  return NULL;
}

static int FN(run)(FN(Job) * j, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  j->next_tile = 0;
  if (threads == 1) {
    FN(worker)(j);
    return 0;
  }
  pthread_t th[256];
  int started = 0;
  for (int i = 0; i < threads; ++i) {
    if (pthread_create(&th[i], NULL, FN(worker), j) != 0) break;
    ++started;
  }
  if (started == 0) FN(worker)(j);
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  return 0;
}

int FN(ro_rasterize_fwd)(int C, int N, int D, int W, int H, int tile_w, int tile_h, const REAL* means2d,
                         const REAL* conics, const REAL* colors, const REAL* opac, const REAL* ray_ts,
                         const REAL* ray_planes, const REAL* normals, const REAL* Ks, const REAL* backgrounds,
                         const int32_t* offsets, const int32_t* flatten_ids, int64_t M, REAL* out_colors,
                         REAL* out_alphas, REAL* out_dexp, REAL* out_dmed, REAL* out_normals, int32_t* last_ids,
                         int32_t* median_ids, uint8_t* fragile, int64_t* counters, double fragile_margin_scale,
                         int threads) {
  FN(Job) j;
  memset(&j, 0, sizeof(j));
  j.C = C; j.N = N; j.D = D; j.W = W; j.H = H; j.tile_w = tile_w; j.tile_h = tile_h;
  j.means2d = means2d; j.conics = conics; j.colors = colors; j.opac = opac; j.ray_ts = ray_ts;
  j.ray_planes = ray_planes; j.normals = normals; j.Ks = Ks; j.backgrounds = backgrounds;
  j.offsets = offsets; j.flatten_ids = flatten_ids; j.M = M;
  j.out_colors = out_colors; j.out_alphas = out_alphas; j.out_dexp = out_dexp; j.out_dmed = out_dmed;
  j.out_normals = out_normals; j.last_ids = last_ids; j.median_ids = median_ids; j.fragile = fragile;
  j.counters = counters;
  j.margin = fragile_margin_scale > 0 ? fragile_margin_scale : 1.0;
  j.backward = 0;
  return FN(run)(&j, threads);
}

int FN(ro_rasterize_bwd)(int C, int N, int D, int W, int H, int tile_w, int tile_h, const REAL* means2d,
                         const REAL* conics, const REAL* colors, const REAL* opac, const REAL* ray_ts,
                         const REAL* ray_planes, const REAL* normals, const REAL* Ks, const REAL* backgrounds,
                         const int32_t* offsets, const int32_t* flatten_ids, int64_t M, const int32_t* last_ids,
                         const int32_t* median_ids, const REAL* v_colors, const REAL* v_alphas, const REAL* v_dexp,
                         const REAL* v_dmed, const REAL* v_normals, REAL* g_means2d, REAL* g_conics, REAL* g_colors,
                         REAL* g_opac, REAL* g_ray_ts, REAL* g_ray_planes, REAL* g_normals, REAL* g_backgrounds,
                         int threads) {
  FN(Job) j;
  memset(&j, 0, sizeof(j));
  j.C = C; j.N = N; j.D = D; j.W = W; j.H = H; j.tile_w = tile_w; j.tile_h = tile_h;
  j.means2d = means2d; j.conics = conics; j.colors = colors; j.opac = opac; j.ray_ts = ray_ts;
  j.ray_planes = ray_planes; j.normals = normals; j.Ks = Ks; j.backgrounds = backgrounds;
  j.offsets = offsets; j.flatten_ids = flatten_ids; j.M = M;
  j.last_ids = (int32_t*)last_ids; j.median_ids = (int32_t*)median_ids;
  j.v_colors = v_colors; j.v_alphas = v_alphas; j.v_dexp = v_dexp; j.v_dmed = v_dmed; j.v_normals = v_normals;
  j.g_means2d = g_means2d; j.g_conics = g_conics; j.g_colors = g_colors; j.g_opac = g_opac; j.g_ray_ts = g_ray_ts;
  j.g_ray_planes = g_ray_planes; j.g_normals = g_normals; j.g_backgrounds = g_backgrounds;
  j.backward = 1;
  return FN(run)(&j, threads);
}

#undef FN
#undef CAT
#undef CAT2
