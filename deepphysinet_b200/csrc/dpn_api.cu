// extern "C" surface of libdpn_b200.so (include/dpn_b200.h): argument validation, mode dispatch.
#include <stdarg.h>

#include <atomic>

#include "dpn_fp32.cuh"
#include "dpn_tc.cuh"

namespace dpn {

int run_sampler(const DpnSampler& S, const float* coarse, const float* x, const float* y, const float* t,
                float* coord_data, float* f, cudaStream_t st);

int run_query_gen(const DpnQueryGen& G, const DpnSampler* S, const float* coarse, float* x, float* y, float* t,
                  float* coord_data, float* f, cudaStream_t st);

static thread_local char g_err[1024] = "";
thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int f32_chunk(const DpnShape& s) {
  int c = s.chunk > 0 ? s.chunk : f32::DEFAULT_CHUNK;
  return c < s.N ? c : s.N;
}

static int planes(const DpnShape& s) { return (s.mode == DPN_MODE_BF16X3 || s.mode == DPN_MODE_F16X3 || s.mode == DPN_MODE_F16X3A) ? 2 : 1; }

static int tc_chunk(const DpnShape& s) {
  int c = s.chunk > 0 ? (s.chunk + 127) / 128 * 128 : tc::default_chunk(s.B, planes(s));
  const int n = (s.N + 127) / 128 * 128;
  return c < n ? c : n;
}

static int check_shape(const DpnShape* s) {
  if (!s) { set_error("shape is NULL"); return DPN_E_INVALID; }
  if (s->B <= 0 || s->N <= 0 || s->K <= 0 || s->K > DPN_MAX_NETS) {
    set_error("bad shape B=%d N=%d K=%d (need B>0, N>0, 1<=K<=6)", s->B, s->N, s->K);
    return DPN_E_INVALID;
  }
  if (s->mode < DPN_MODE_FP32 || s->mode > DPN_MODE_F16X3A) {
    set_error("unknown mode %d", s->mode);
    return DPN_E_INVALID;
  }
  return 0;
}

static int check_device() {
  // cudaGetDeviceProperties costs milliseconds per call: ask for the one attribute, once per device
  static std::atomic<int> major_of[64];                               // zero-initialised; threads (one per GPU) may race to fill an entry
  int dev = 0;
  DPN_CUDA_OK(cudaGetDevice(&dev));
  int major = (dev >= 0 && dev < 64) ? major_of[dev].load(std::memory_order_relaxed) : 0;
  if (major == 0) {
    DPN_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (dev >= 0 && dev < 64) major_of[dev].store(major, std::memory_order_relaxed);
  }
  if (major != 10) {
    set_error("libdpn_b200 is built for sm_100a only; device %d has compute capability major %d", dev, major);
    return DPN_E_UNSUPPORTED;
  }
  return 0;
}

static size_t ws_bytes(const DpnShape& s) {
  if (s.mode == DPN_MODE_FP32) return f32::workspace_bytes(f32_chunk(s), s.K, s.B);
  return tc::workspace_bytes(tc_chunk(s), s.K, s.B, planes(s));
}

static int dispatch(Job& job, void* stream) {
  g_launches = 0;
  int rc = check_device();
  if (rc) return rc;
  const size_t need = ws_bytes(job.shape);
  if (!job.workspace || job.workspace_bytes < need) {
    set_error("workspace too small: have %zu bytes, need %zu", job.workspace_bytes, need);
    return DPN_E_WORKSPACE;
  }
  if ((reinterpret_cast<uintptr_t>(job.workspace) & 255) != 0) {
    set_error("workspace must be 256-byte aligned");
    return DPN_E_INVALID;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // `mode` is honoured on every surface: pre-encoded coordinates (coord_pe), an explicit skip (ref) and K < 6 nets run on the
  // tensor cores as well (tc::encode_kernel takes the caller's encoding; the epilogues take ref)
  const bool tensor = job.shape.mode != DPN_MODE_FP32;
  job.chunk = tensor ? tc_chunk(job.shape) : f32_chunk(job.shape);
  return tensor ? tc::run(job, st) : f32::run(job, st);
}

static int check_common(const DpnShape* s, const DpnConsts* c, const DpnPoints* p, const DpnWeights* w) {
  int rc = check_shape(s);
  if (rc) return rc;
  if (!c || !p || !w) { set_error("consts / points / weights is NULL"); return DPN_E_INVALID; }
  const void* ws[] = {w->W1, w->b1, w->W2, w->b2, w->e, w->Wd, w->bd, w->Wa, w->ba, w->Wb, w->bb, w->wo, w->bo};
  for (const void* q : ws)
    if (!q || (reinterpret_cast<uintptr_t>(q) & 15)) { set_error("a weight pointer is NULL or not 16-byte aligned"); return DPN_E_INVALID; }
  if (!p->coord_data) { set_error("coord_data is NULL"); return DPN_E_INVALID; }
  if (!p->coord_pe && !(p->x && p->y && p->t)) { set_error("need either coord_pe or x,y,t"); return DPN_E_INVALID; }
  return 0;
}

static int check_grads(const DpnGrads* g) {
  const void* gs[] = {g->W1, g->b1, g->W2, g->b2, g->e, g->Wd, g->bd, g->Wa, g->ba, g->Wb, g->bb, g->wo, g->bo};
  for (const void* q : gs)
    if (!q || (reinterpret_cast<uintptr_t>(q) & 15)) { set_error("a gradient pointer is NULL or not 16-byte aligned"); return DPN_E_INVALID; }
  return 0;
}

}  // namespace dpn

using namespace dpn;

extern "C" {

int dpn_abi_version(void) { return DPN_ABI_VERSION; }

size_t dpn_last_error(char* buf, size_t cap) {
  if (!buf || cap == 0) return strlen(g_err);
  strncpy(buf, g_err, cap - 1);
  buf[cap - 1] = 0;
  return strlen(buf);
}

int dpn_last_launch_count(void) { return g_launches; }

int dpn_workspace_bytes(const DpnShape* shape, size_t* bytes) {
  int rc = check_shape(shape);
  if (rc) return rc;
  if (!bytes) { set_error("bytes is NULL"); return DPN_E_INVALID; }
  *bytes = ws_bytes(*shape);
  return 0;
}

int dpn_pde_margin_fwd_bwd(const DpnShape* shape, const DpnConsts* consts, const DpnPoints* pts, const DpnWeights* w,
                           const DpnMargin* margin, const DpnPdeOut* out, const DpnGrads* grads, void* workspace,
                           size_t workspace_bytes, void* cuda_stream) {
  int rc = check_common(shape, consts, pts, w);
  if (rc) return rc;
  if (shape->K != 6) { set_error("the PDE residual needs K = 6 nets (u,v,p,T,q,rho), got %d", shape->K); return DPN_E_INVALID; }
  if (!(pts->x && pts->y && pts->t && pts->f)) { set_error("the PDE path needs x, y, t and f"); return DPN_E_INVALID; }
  if (pts->coord_pe || pts->ref) { set_error("the PDE path derives the encoding from x,y,t and the skip from coord_data"); return DPN_E_INVALID; }
  if (!out || !out->loss_terms) { set_error("out->loss_terms is NULL"); return DPN_E_INVALID; }
  if (grads && (rc = check_grads(grads))) return rc;
  Job job;
  memset(&job, 0, sizeof(job));
  job.kind = JOB_PDE; job.shape = *shape; job.dc = make_dev_consts(*consts);
  if (margin && (!margin->target || !margin->loss || !(margin->beta > 0.0))) {
    set_error("margin needs target, loss and beta > 0");
    return DPN_E_INVALID;
  }
  job.pts = pts; job.w = w; job.out = out; job.grads = grads; job.margin = margin;
  job.workspace = workspace; job.workspace_bytes = workspace_bytes;
  return dispatch(job, cuda_stream);
}

int dpn_pde_fwd_bwd(const DpnShape* shape, const DpnConsts* consts, const DpnPoints* pts, const DpnWeights* w,
                    const DpnPdeOut* out, const DpnGrads* grads, void* workspace, size_t workspace_bytes,
                    void* cuda_stream) {
  return dpn_pde_margin_fwd_bwd(shape, consts, pts, w, nullptr, out, grads, workspace, workspace_bytes, cuda_stream);
}

int dpn_decoder_fwd(const DpnShape* shape, const DpnConsts* consts, const DpnPoints* pts, const DpnWeights* w,
                    float* o, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  int rc = check_common(shape, consts, pts, w);
  if (rc) return rc;
  if (!o) { set_error("o is NULL"); return DPN_E_INVALID; }
  if (!pts->ref && shape->K != 6) { set_error("ref may only be NULL for K = 6"); return DPN_E_INVALID; }
  Job job;
  memset(&job, 0, sizeof(job));
  job.kind = JOB_DEC_FWD; job.shape = *shape; job.dc = make_dev_consts(*consts);
  job.pts = pts; job.w = w; job.o = o;
  job.workspace = workspace; job.workspace_bytes = workspace_bytes;
  return dispatch(job, cuda_stream);
}

int dpn_decoder_bwd(const DpnShape* shape, const DpnConsts* consts, const DpnPoints* pts, const DpnWeights* w,
                    const float* d_o, const DpnGrads* grads, void* workspace, size_t workspace_bytes,
                    void* cuda_stream) {
  int rc = check_common(shape, consts, pts, w);
  if (rc) return rc;
  if (!d_o || !grads) { set_error("d_o / grads is NULL"); return DPN_E_INVALID; }
  if (!pts->ref && shape->K != 6) { set_error("ref may only be NULL for K = 6"); return DPN_E_INVALID; }
  if ((rc = check_grads(grads))) return rc;
  Job job;
  memset(&job, 0, sizeof(job));
  job.kind = JOB_DEC_BWD; job.shape = *shape; job.dc = make_dev_consts(*consts);
  job.pts = pts; job.w = w; job.d_o = d_o; job.grads = grads;
  job.workspace = workspace; job.workspace_bytes = workspace_bytes;
  return dispatch(job, cuda_stream);
}

int dpn_sample_field(const DpnSampler* s, const float* coarse, const float* x, const float* y, const float* t,
                     float* coord_data, float* f, void* cuda_stream) {
  if (!s || !coarse || !x || !y || !t || !coord_data) { set_error("dpn_sample_field: NULL argument"); return DPN_E_INVALID; }
  if (s->B <= 0 || s->N <= 0 || s->Tt < 2 || s->Hc < 2 || s->Wc < 2 || !(s->cells_per_coarse > 0) || !(s->t_step > 0)) {
    set_error("dpn_sample_field: bad sampler shape B=%d N=%d T=%d H=%d W=%d", s->B, s->N, s->Tt, s->Hc, s->Wc);
    return DPN_E_INVALID;
  }
  g_launches = 0;
  int rc = check_device();
  if (rc) return rc;
  return run_sampler(*s, coarse, x, y, t, coord_data, f, reinterpret_cast<cudaStream_t>(cuda_stream));
}

int dpn_generate_queries(const DpnQueryGen* g, const DpnSampler* s, const float* coarse, float* x, float* y, float* t,
                         float* coord_data, float* f, void* cuda_stream) {
  if (!g || !x || !y || !t) { set_error("dpn_generate_queries: NULL argument"); return DPN_E_INVALID; }
  if (g->B <= 0 || g->N <= 0 || g->lat_size < 2 || g->lon_size < 2 || g->t_steps < 1) {
    set_error("dpn_generate_queries: bad generator B=%d N=%d grid %dx%d t_steps=%d", g->B, g->N, g->lat_size, g->lon_size, g->t_steps);
    return DPN_E_INVALID;
  }
  if (s) {
    if (!coarse || !coord_data) { set_error("dpn_generate_queries: sampler given without coarse / coord_data"); return DPN_E_INVALID; }
    if (s->B != g->B || s->N != g->N || s->Tt < 2 || s->Hc < 2 || s->Wc < 2 || !(s->cells_per_coarse > 0) || !(s->t_step > 0)) {
      set_error("dpn_generate_queries: sampler shape does not match the generator");
      return DPN_E_INVALID;
    }
  }
  g_launches = 0;
  int rc = check_device();
  if (rc) return rc;
  return run_query_gen(*g, s, coarse, x, y, t, coord_data, f, reinterpret_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
