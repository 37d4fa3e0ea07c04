// capi.cu -- host side of libesfm_match.so: the C ABI declared in include/esfm_match.h.
//
// Host logic only (bank packing, pair-batch chunking, scratch management, result bookkeeping); every
// distance, selection, ratio, cross-check and compaction step runs in the CUDA kernels of this library.
// There is no CPU fallback: if the device or a kernel fails the call fails.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "tc_layout.cuh"

using namespace esfm;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess)                                                                                 \
            return fail(ESFM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------------
struct esfm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    bool profiling = true;
    int tc_qtiles = 1;                     // TC sweep geometry, SURF: query tiles per block (1 or 2; $ESFM_TC_QT)
    int tc_qtiles_orb = 1;                 // TC sweep geometry, ORB (1 or 2; $ESFM_TC_QT_ORB): 3 accumulator stages either way
    int surf_bf = 0;                       // SURF tensor-core sweep with the branch-free row selection (EXPERIMENTAL, $ESFM_TC_SURF_BF=1)
    int orb_z = 1;                         // ORB tensor-core sweep with the "Z" operand encoding (packed keys from the MMA): default;
                                           // $ESFM_ORB_Z=0 selects the +-1 encoding with the generic epilogue
    int hamming_engine = ESFM_HAMMING_ENGINE_TC;     // which sweep kernel serves ESFM_KIND_B256 (esfm_set_hamming_engine / $ESFM_HAMMING_ENGINE)
    int l2_engine = ESFM_L2_ENGINE_TC;     // which sweep kernel serves ESFM_KIND_F32X64 (esfm_set_l2_engine / $ESFM_L2_ENGINE)
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    esfm_stats_t stats{};
    // device scratch, grown on demand
    u64* keys = nullptr;           size_t keys_bytes = 0;
    uint32_t* col_thr = nullptr;   size_t col_thr_elems = 0;
    esfm_dmatch_t* arena = nullptr; size_t arena_cap = 0;   // in matches
    PairDesc* d_pairs = nullptr;   size_t pairs_cap = 0;
    unsigned long long* d_pair_off = nullptr;
    int32_t* d_pair_cnt = nullptr;
    unsigned long long* d_cursor = nullptr;  // [0] cursor, [1] overflow flag (as int)
    // pinned host staging (small metadata) + a pool of large pinned buffers that banks (upload staging) and
    // results (downloaded matches) borrow, so steady-state calls never allocate or zero-fill host memory
    void* h_stage = nullptr;       size_t h_stage_bytes = 0;
    struct Pinned { void* ptr; size_t bytes; bool in_use; };
    std::vector<Pinned> pool;
    uint64_t arena_generation = 0;
};

static void* pool_acquire(esfm_ctx* ctx, size_t bytes, size_t* got) {
    size_t best = (size_t)-1;
    for (size_t i = 0; i < ctx->pool.size(); ++i)
        if (!ctx->pool[i].in_use && ctx->pool[i].bytes >= bytes && (best == (size_t)-1 || ctx->pool[i].bytes < ctx->pool[best].bytes)) best = i;
    if (best != (size_t)-1) {
        ctx->pool[best].in_use = true;
        if (got) *got = ctx->pool[best].bytes;
        return ctx->pool[best].ptr;
    }
    // drop idle buffers that were too small, then allocate with 25% headroom
    for (size_t i = 0; i < ctx->pool.size();) {
        if (!ctx->pool[i].in_use) { cudaFreeHost(ctx->pool[i].ptr); ctx->pool.erase(ctx->pool.begin() + i); } else ++i;
    }
    size_t want = bytes + bytes / 4 + 4096;
    void* p = nullptr;
    if (cudaMallocHost(&p, want) != cudaSuccess) {
        want = bytes;
        if (cudaMallocHost(&p, want) != cudaSuccess) return nullptr;
    }
    ctx->pool.push_back({p, want, true});
    if (got) *got = want;
    return p;
}

static void pool_release(esfm_ctx* ctx, void* ptr) {
    if (!ptr) return;
    for (auto& b : ctx->pool)
        if (b.ptr == ptr) { b.in_use = false; return; }
}


struct esfm_bank {
    esfm_ctx* ctx = nullptr;
    int kind = 0;
    int n_frames = 0;
    std::vector<int> rows;                    // per frame, -1 = not set
    uint8_t* h_up = nullptr;                  // pinned upload staging (borrowed from the ctx pool until commit)
    size_t h_up_cap = 0, h_up_used = 0;
    std::vector<size_t> host_off;             // per frame offset into h_up ((size_t)-1 = no host data)
    std::vector<const void*> host_ext;        // per frame caller-owned pinned source (esfm_bank_set_frame_pinned), else nullptr
    bool committed = false;
    bool device_allocated = false;
    // device
    void* d_rows = nullptr;  size_t rows_bytes = 0;
    float* d_kmajor = nullptr; size_t kmajor_bytes = 0;
    unsigned char* d_tc = nullptr; size_t tc_bytes = 0;   // tensor-core operand images (built on first use), main then aug
    int tc_z = 0;                                         // ... of a B256 bank: 0 = +-1 encoding, 1 = "Z" encoding (tc_layout.cuh)
    int* d_frame_rows = nullptr;
    int* d_row_off = nullptr;
    int* d_tile_off = nullptr;
    // host copies
    std::vector<int> row_off, tile_off;
    int max_rows = 0;
    size_t row_bytes() const { return kind == ESFM_KIND_F32X64 ? kDim * sizeof(float) : 32; }
};

struct esfm_results {
    esfm_ctx* ctx = nullptr;
    std::vector<PairDesc> pairs;
    std::vector<int32_t> counts;
    std::vector<uint64_t> offsets;          // segment index << 40 | offset (in matches) inside that segment
    struct Segment { esfm_dmatch_t* ptr; size_t count; };
    std::vector<Segment> segments;          // one pinned buffer per chunk, borrowed from the ctx pool
    int kind = -1, cross_check = 0;         // what the batch was matched with (kept in the match file)
    double ratio = 0.0;
    bool heap_segments = false;             // esfm_results_load: segments are plain malloc memory, there is no ctx
    std::unordered_map<uint64_t, int64_t> index;
    bool fetched = true;
    // device-resident variant (single chunk only)
    uint64_t arena_generation = 0;
    uint64_t device_matches = 0;
    int64_t total_matches = 0;
};

static int set_device(esfm_ctx* ctx) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    return ESFM_OK;
}

template <typename T>
static int grow(T** ptr, size_t* cap_elems, size_t need_elems) {
    if (*cap_elems >= need_elems && *ptr) return ESFM_OK;
    if (*ptr) CUDA_TRY(cudaFree(*ptr));
    *ptr = nullptr;
    *cap_elems = 0;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, need_elems * sizeof(T));
    if (e != cudaSuccess) return fail(ESFM_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", need_elems * sizeof(T), cudaGetErrorString(e));
    *ptr = (T*)p;
    *cap_elems = need_elems;
    return ESFM_OK;
}

static int grow_stage(esfm_ctx* ctx, size_t bytes) {
    if (ctx->h_stage_bytes >= bytes) return ESFM_OK;
    if (ctx->h_stage) CUDA_TRY(cudaFreeHost(ctx->h_stage));
    ctx->h_stage = nullptr;
    ctx->h_stage_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 20);
    cudaError_t e = cudaMallocHost(&ctx->h_stage, want);
    if (e != cudaSuccess) return fail(ESFM_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
    ctx->h_stage_bytes = want;
    return ESFM_OK;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int esfm_abi_version(void) { return ESFM_ABI_VERSION; }
extern "C" const char* esfm_last_error(void) { return g_last_error.c_str(); }

extern "C" int esfm_init(int device, void* cuda_stream, esfm_ctx_t** out) {
    if (!out) return fail(ESFM_ERR_INVALID, "esfm_init: ctx is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(ESFM_ERR_CUDA, "esfm_init: no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(ESFM_ERR_INVALID, "esfm_init: device %d out of range [0,%d)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(ESFM_ERR_CUDA, "esfm_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    esfm_ctx* ctx = new (std::nothrow) esfm_ctx();
    if (!ctx) return fail(ESFM_ERR_NOMEM, "esfm_init: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* qt = getenv("ESFM_TC_QT")) ctx->tc_qtiles = atoi(qt) == 2 ? 2 : 1;
    if (const char* qt = getenv("ESFM_TC_QT_ORB")) ctx->tc_qtiles_orb = atoi(qt) == 2 ? 2 : 1;
    if (const char* z = getenv("ESFM_ORB_Z")) ctx->orb_z = atoi(z) != 0;
    if (const char* bf = getenv("ESFM_TC_SURF_BF")) ctx->surf_bf = atoi(bf) != 0;
    if (const char* eng = getenv("ESFM_HAMMING_ENGINE")) {
        if (!strcmp(eng, "tc") || !strcmp(eng, "tensor")) ctx->hamming_engine = ESFM_HAMMING_ENGINE_TC;
        else if (!strcmp(eng, "popc")) ctx->hamming_engine = ESFM_HAMMING_ENGINE_POPC;
        else { delete ctx; return fail(ESFM_ERR_INVALID, "ESFM_HAMMING_ENGINE=%s: expected 'popc' or 'tc'", eng); }
    }
    if (const char* eng = getenv("ESFM_L2_ENGINE")) {
        if (!strcmp(eng, "tc") || !strcmp(eng, "tensor")) ctx->l2_engine = ESFM_L2_ENGINE_TC;
        else if (!strcmp(eng, "ffma")) ctx->l2_engine = ESFM_L2_ENGINE_FFMA;
        else { delete ctx; return fail(ESFM_ERR_INVALID, "ESFM_L2_ENGINE=%s: expected 'ffma' or 'tc'", eng); }
    }
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return fail(ESFM_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e)); }
        ctx->own_stream = true;
    }
    for (auto& ev : ctx->ev) {
        e = cudaEventCreate(&ev);
        if (e != cudaSuccess) { delete ctx; return fail(ESFM_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(e)); }
    }
    e = cudaMalloc((void**)&ctx->d_cursor, 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) { delete ctx; return fail(ESFM_ERR_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    {   // banks come and go (one per all-pairs call in the per-call API): allocate them stream-ordered from the
        // device's memory pool and keep freed blocks cached, so steady state never pays cudaMalloc / cudaFree
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    *out = ctx;
    return ESFM_OK;
}

extern "C" int esfm_destroy(esfm_ctx_t* ctx) {
    if (!ctx) return ESFM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->keys); cudaFree(ctx->col_thr); cudaFree(ctx->arena); cudaFree(ctx->d_pairs); cudaFree(ctx->d_pair_off);
    cudaFree(ctx->d_pair_cnt); cudaFree(ctx->d_cursor);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (auto& b : ctx->pool) cudaFreeHost(b.ptr);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ESFM_OK;
}

extern "C" int esfm_synchronize(esfm_ctx_t* ctx) {
    if (!ctx) return fail(ESFM_ERR_INVALID, "ctx is NULL");
    if (int rc = set_device(ctx)) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ESFM_OK;
}

extern "C" int esfm_get_stats(esfm_ctx_t* ctx, esfm_stats_t* out) {
    if (!ctx || !out) return fail(ESFM_ERR_INVALID, "esfm_get_stats: NULL argument");
    *out = ctx->stats;
    return ESFM_OK;
}

extern "C" int esfm_set_profiling(esfm_ctx_t* ctx, int enabled) {
    if (!ctx) return fail(ESFM_ERR_INVALID, "ctx is NULL");
    ctx->profiling = enabled != 0;
    return ESFM_OK;
}

extern "C" int esfm_set_l2_engine(esfm_ctx_t* ctx, int engine) {
    if (!ctx) return fail(ESFM_ERR_INVALID, "ctx is NULL");
    if (engine != ESFM_L2_ENGINE_FFMA && engine != ESFM_L2_ENGINE_TC) return fail(ESFM_ERR_INVALID, "unknown L2 engine %d", engine);
    ctx->l2_engine = engine;
    return ESFM_OK;
}

extern "C" int esfm_set_hamming_engine(esfm_ctx_t* ctx, int engine) {
    if (!ctx) return fail(ESFM_ERR_INVALID, "ctx is NULL");
    if (engine != ESFM_HAMMING_ENGINE_POPC && engine != ESFM_HAMMING_ENGINE_TC) return fail(ESFM_ERR_INVALID, "unknown Hamming engine %d", engine);
    ctx->hamming_engine = engine;
    return ESFM_OK;
}

extern "C" int esfm_get_hamming_engine(esfm_ctx_t* ctx, int* engine) {
    if (!ctx || !engine) return fail(ESFM_ERR_INVALID, "esfm_get_hamming_engine: NULL argument");
    *engine = ctx->hamming_engine;
    return ESFM_OK;
}

extern "C" int esfm_get_l2_engine(esfm_ctx_t* ctx, int* engine) {
    if (!ctx || !engine) return fail(ESFM_ERR_INVALID, "esfm_get_l2_engine: NULL argument");
    *engine = ctx->l2_engine;
    return ESFM_OK;
}

extern "C" int esfm_device_sm_count(esfm_ctx_t* ctx, int* sms) {
    if (!ctx || !sms) return fail(ESFM_ERR_INVALID, "esfm_device_sm_count: NULL argument");
    *sms = ctx->sm_count;
    return ESFM_OK;
}

// ------------------------------------------------------------------------------------------------
// bank
// ------------------------------------------------------------------------------------------------
extern "C" int esfm_bank_create(esfm_ctx_t* ctx, esfm_kind kind, int n_frames, esfm_bank_t** out) {
    if (!ctx || !out) return fail(ESFM_ERR_INVALID, "esfm_bank_create: NULL argument");
    *out = nullptr;
    if (kind != ESFM_KIND_F32X64 && kind != ESFM_KIND_B256) return fail(ESFM_ERR_INVALID, "esfm_bank_create: unknown kind %d", (int)kind);
    if (n_frames < 0) return fail(ESFM_ERR_INVALID, "esfm_bank_create: n_frames < 0");
    esfm_bank* b = new (std::nothrow) esfm_bank();
    if (!b) return fail(ESFM_ERR_NOMEM, "out of host memory");
    b->ctx = ctx;
    b->kind = kind;
    b->n_frames = n_frames;
    b->rows.assign(n_frames, -1);
    b->host_off.assign(n_frames, (size_t)-1);
    b->host_ext.assign(n_frames, nullptr);
    *out = b;
    return ESFM_OK;
}

static int check_frame_limits(esfm_bank* b, int rows) {
    const int lim = b->kind == ESFM_KIND_F32X64 ? sweep_l2_max_rows() : sweep_hamming_max_rows();
    if (rows > lim) return fail(ESFM_ERR_CAPACITY, "frame has %d rows; this build supports at most %d per frame for kind %d", rows, lim, b->kind);
    return ESFM_OK;
}

extern "C" int esfm_bank_set_frame(esfm_bank_t* b, int frame_id, const void* data, int rows, int cols, size_t step_bytes) {
    if (!b) return fail(ESFM_ERR_INVALID, "bank is NULL");
    if (b->committed || b->device_allocated) return fail(ESFM_ERR_STATE, "esfm_bank_set_frame: bank already committed");
    if (frame_id < 0 || frame_id >= b->n_frames) return fail(ESFM_ERR_INVALID, "frame_id %d out of range [0,%d)", frame_id, b->n_frames);
    if (rows < 0) return fail(ESFM_ERR_INVALID, "rows < 0");
    const int want_cols = b->kind == ESFM_KIND_F32X64 ? kDim : 32;
    if (rows > 0 && cols != want_cols)
        return fail(ESFM_ERR_INVALID, "kind %d needs %d columns per descriptor, got %d", b->kind, want_cols, cols);
    const size_t rb = b->row_bytes();
    if (rows > 0 && !data) return fail(ESFM_ERR_INVALID, "data is NULL with rows > 0");
    if (rows > 0 && step_bytes < rb) return fail(ESFM_ERR_INVALID, "step_bytes %zu smaller than a row (%zu)", step_bytes, rb);
    if (int rc = check_frame_limits(b, rows)) return rc;
    // append to the pinned upload staging (frames set in order land exactly where the single H2D copy wants them)
    const size_t need = (size_t)rows * rb;
    if (b->h_up_used + need > b->h_up_cap) {
        size_t cap = 0;
        const size_t want = std::max((b->h_up_used + need) * 2, (size_t)8 << 20);
        uint8_t* nb = (uint8_t*)pool_acquire(b->ctx, want, &cap);
        if (!nb) return fail(ESFM_ERR_NOMEM, "pinned staging allocation of %zu bytes failed", want);
        if (b->h_up_used) memcpy(nb, b->h_up, b->h_up_used);
        pool_release(b->ctx, b->h_up);
        b->h_up = nb;
        b->h_up_cap = cap;
    }
    uint8_t* dst = b->h_up + b->h_up_used;
    if (step_bytes == rb) {
        if (need) memcpy(dst, data, need);
    } else {
        for (int r = 0; r < rows; ++r) memcpy(dst + (size_t)r * rb, (const uint8_t*)data + (size_t)r * step_bytes, rb);
    }
    b->host_off[frame_id] = b->h_up_used;
    b->host_ext[frame_id] = nullptr;
    b->h_up_used += need;
    b->rows[frame_id] = rows;
    return ESFM_OK;
}

extern "C" int esfm_bank_set_frame_pinned(esfm_bank_t* b, int frame_id, const void* data, int rows, int cols, size_t step_bytes) {
    if (!b) return fail(ESFM_ERR_INVALID, "bank is NULL");
    if (b->committed || b->device_allocated) return fail(ESFM_ERR_STATE, "esfm_bank_set_frame_pinned: bank already committed");
    if (frame_id < 0 || frame_id >= b->n_frames) return fail(ESFM_ERR_INVALID, "frame_id %d out of range [0,%d)", frame_id, b->n_frames);
    if (rows < 0) return fail(ESFM_ERR_INVALID, "rows < 0");
    const int want_cols = b->kind == ESFM_KIND_F32X64 ? kDim : 32;
    if (rows > 0 && cols != want_cols)
        return fail(ESFM_ERR_INVALID, "kind %d needs %d columns per descriptor, got %d", b->kind, want_cols, cols);
    if (rows > 0 && !data) return fail(ESFM_ERR_INVALID, "data is NULL with rows > 0");
    if (rows > 0 && step_bytes != b->row_bytes())
        return fail(ESFM_ERR_INVALID, "esfm_bank_set_frame_pinned needs densely packed rows (step_bytes %zu != %zu)", step_bytes, b->row_bytes());
    if (int rc = check_frame_limits(b, rows)) return rc;
    if (rows > 0) {
        if (int rc = set_device(b->ctx)) return rc;
        cudaPointerAttributes at{};
        const cudaError_t e = cudaPointerGetAttributes(&at, data);
        if (e != cudaSuccess || at.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            return fail(ESFM_ERR_INVALID, "esfm_bank_set_frame_pinned: frame %d is not in page-locked host memory", frame_id);
        }
    }
    b->host_off[frame_id] = (size_t)-1;
    b->host_ext[frame_id] = rows > 0 ? data : nullptr;
    b->rows[frame_id] = rows;
    return ESFM_OK;
}

extern "C" int esfm_bank_set_frame_rows(esfm_bank_t* b, int frame_id, int rows) {
    if (!b) return fail(ESFM_ERR_INVALID, "bank is NULL");
    if (b->committed || b->device_allocated) return fail(ESFM_ERR_STATE, "bank already committed");
    if (frame_id < 0 || frame_id >= b->n_frames) return fail(ESFM_ERR_INVALID, "frame_id %d out of range", frame_id);
    if (rows < 0) return fail(ESFM_ERR_INVALID, "rows < 0");
    if (int rc = check_frame_limits(b, rows)) return rc;
    b->rows[frame_id] = rows;
    b->host_off[frame_id] = (size_t)-1;
    b->host_ext[frame_id] = nullptr;
    return ESFM_OK;
}

static int bank_alloc_layout(esfm_bank* b) {
    esfm_ctx* ctx = b->ctx;
    if (int rc = set_device(ctx)) return rc;
    for (int f = 0; f < b->n_frames; ++f)
        if (b->rows[f] < 0) return fail(ESFM_ERR_STATE, "frame %d was never set", f);
    b->row_off.assign(b->n_frames + 1, 0);
    b->tile_off.assign(b->n_frames + 1, 0);
    b->max_rows = 0;
    for (int f = 0; f < b->n_frames; ++f) {
        b->row_off[f + 1] = b->row_off[f] + b->rows[f];
        b->tile_off[f + 1] = b->tile_off[f] + (b->rows[f] + kTile - 1) / kTile;
        b->max_rows = std::max(b->max_rows, b->rows[f]);
        if (b->row_off[f + 1] < b->row_off[f]) return fail(ESFM_ERR_CAPACITY, "bank exceeds 2^31 rows");
    }
    const size_t total_rows = (size_t)b->row_off[b->n_frames];
    // one extra tile of slack so tile-granular reads past the last frame stay inside the allocation
    b->rows_bytes = (total_rows + kHamTile) * b->row_bytes();
    cudaError_t e = cudaMallocAsync(&b->d_rows, b->rows_bytes, ctx->stream);
    if (e != cudaSuccess) return fail(ESFM_ERR_NOMEM, "cudaMallocAsync(%zu) for the descriptor bank failed: %s", b->rows_bytes, cudaGetErrorString(e));
    // only the slack past the last frame needs defined contents
    CUDA_TRY(cudaMemsetAsync((uint8_t*)b->d_rows + total_rows * b->row_bytes(), 0, (size_t)kHamTile * b->row_bytes(), ctx->stream));
    if (b->kind == ESFM_KIND_F32X64) {
        b->kmajor_bytes = ((size_t)b->tile_off[b->n_frames] + 1) * kTileBytes;
        e = cudaMallocAsync((void**)&b->d_kmajor, b->kmajor_bytes, ctx->stream);
        if (e != cudaSuccess) return fail(ESFM_ERR_NOMEM, "cudaMallocAsync(%zu) for the k-major bank failed: %s", b->kmajor_bytes, cudaGetErrorString(e));
    }
    const size_t nb = (size_t)(b->n_frames + 1) * sizeof(int);
    CUDA_TRY(cudaMallocAsync((void**)&b->d_frame_rows, nb, ctx->stream));
    CUDA_TRY(cudaMallocAsync((void**)&b->d_row_off, nb, ctx->stream));
    CUDA_TRY(cudaMallocAsync((void**)&b->d_tile_off, nb, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(b->d_frame_rows, b->rows.data(), (size_t)b->n_frames * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(b->d_row_off, b->row_off.data(), nb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(b->d_tile_off, b->tile_off.data(), nb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += 3 * nb;
    b->device_allocated = true;
    return ESFM_OK;
}

static int bank_build_derived(esfm_bank* b) {
    esfm_ctx* ctx = b->ctx;
    if (b->kind == ESFM_KIND_F32X64) {
        const int n_tiles = b->tile_off[b->n_frames];
        cudaError_t e = launch_pack_f32((const float*)b->d_rows, b->d_frame_rows, b->d_row_off, b->d_tile_off, b->n_frames,
                                        n_tiles, b->d_kmajor, ctx->stream);
        if (e != cudaSuccess) return fail(ESFM_ERR_CUDA, "pack_f32 kernel launch failed: %s", cudaGetErrorString(e));
        if (n_tiles > 0) ctx->stats.kernel_launches += 1;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    b->committed = true;
    return ESFM_OK;
}

extern "C" int esfm_bank_alloc_device(esfm_bank_t* b) {
    if (!b) return fail(ESFM_ERR_INVALID, "bank is NULL");
    if (b->device_allocated) return fail(ESFM_ERR_STATE, "bank already allocated");
    return bank_alloc_layout(b);
}

extern "C" int esfm_bank_commit(esfm_bank_t* b) {
    if (!b) return fail(ESFM_ERR_INVALID, "bank is NULL");
    if (b->committed) return fail(ESFM_ERR_STATE, "bank already committed");
    if (!b->device_allocated) {
        for (int f = 0; f < b->n_frames; ++f)
            if (b->rows[f] > 0 && b->host_off[f] == (size_t)-1 && !b->host_ext[f])
                return fail(ESFM_ERR_STATE, "frame %d has rows declared but no host data; use esfm_bank_alloc_device + esfm_bank_commit_device", f);
        if (int rc = bank_alloc_layout(b)) return rc;
    }
    esfm_ctx* ctx = b->ctx;
    const size_t rb = b->row_bytes();
    const size_t total = (size_t)b->row_off[b->n_frames] * rb;
    if (total > 0) {
        bool in_order = b->h_up_used == total;
        for (int f = 0; f < b->n_frames && in_order; ++f)
            if (b->rows[f] > 0 && (b->host_ext[f] || b->host_off[f] != (size_t)b->row_off[f] * rb)) in_order = false;
        if (in_order) {   // one pinned -> device copy
            CUDA_TRY(cudaMemcpyAsync(b->d_rows, b->h_up, total, cudaMemcpyHostToDevice, ctx->stream));
        } else {          // frames were set out of order (or re-set): one copy per frame
            for (int f = 0; f < b->n_frames; ++f)
                if (b->rows[f] > 0)     // caller-owned pinned source (no staging copy was made) or the staging buffer
                    CUDA_TRY(cudaMemcpyAsync((uint8_t*)b->d_rows + (size_t)b->row_off[f] * rb,
                                             b->host_ext[f] ? (const uint8_t*)b->host_ext[f] : b->h_up + b->host_off[f],
                                             (size_t)b->rows[f] * rb, cudaMemcpyHostToDevice, ctx->stream));
        }
        ctx->stats.h2d_bytes += total;
    }
    const int rc = bank_build_derived(b);   // synchronises the stream: the staging buffer is free again
    pool_release(ctx, b->h_up);
    b->h_up = nullptr;
    b->h_up_cap = b->h_up_used = 0;
    return rc;
}

extern "C" int esfm_bank_device_rows(esfm_bank_t* b, void** dev_ptr, size_t* bytes) {
    if (!b || !dev_ptr || !bytes) return fail(ESFM_ERR_INVALID, "esfm_bank_device_rows: NULL argument");
    if (!b->device_allocated) return fail(ESFM_ERR_STATE, "bank has no device storage yet");
    *dev_ptr = b->d_rows;
    *bytes = (size_t)b->row_off[b->n_frames] * b->row_bytes();
    return ESFM_OK;
}

extern "C" int esfm_bank_commit_device(esfm_bank_t* b) {
    if (!b) return fail(ESFM_ERR_INVALID, "bank is NULL");
    if (!b->device_allocated) return fail(ESFM_ERR_STATE, "call esfm_bank_alloc_device first");
    if (int rc = set_device(b->ctx)) return rc;
    return bank_build_derived(b);
}

extern "C" int esfm_bank_n_frames(esfm_bank_t* b, int* n) {
    if (!b || !n) return fail(ESFM_ERR_INVALID, "NULL argument");
    *n = b->n_frames;
    return ESFM_OK;
}

extern "C" int esfm_bank_frame_rows(esfm_bank_t* b, int frame_id, int* rows) {
    if (!b || !rows) return fail(ESFM_ERR_INVALID, "NULL argument");
    if (frame_id < 0 || frame_id >= b->n_frames) return fail(ESFM_ERR_INVALID, "frame_id %d out of range", frame_id);
    *rows = b->rows[frame_id];
    return ESFM_OK;
}

extern "C" int esfm_bank_device_bytes(esfm_bank_t* b, size_t* bytes) {
    if (!b || !bytes) return fail(ESFM_ERR_INVALID, "NULL argument");
    *bytes = b->device_allocated ? b->rows_bytes + b->kmajor_bytes + b->tc_bytes : 0;
    return ESFM_OK;
}

extern "C" int esfm_bank_destroy(esfm_bank_t* b) {
    if (!b) return ESFM_OK;
    if (b->ctx) {
        cudaSetDevice(b->ctx->device);
        cudaStream_t s = b->ctx->stream;   // stream-ordered frees: queued behind any work still using the bank
        if (b->d_rows) cudaFreeAsync(b->d_rows, s);
        if (b->d_kmajor) cudaFreeAsync(b->d_kmajor, s);
        if (b->d_tc) cudaFreeAsync(b->d_tc, s);
        if (b->d_frame_rows) cudaFreeAsync(b->d_frame_rows, s);
        if (b->d_row_off) cudaFreeAsync(b->d_row_off, s);
        if (b->d_tile_off) cudaFreeAsync(b->d_tile_off, s);
        pool_release(b->ctx, b->h_up);
    }
    delete b;
    return ESFM_OK;
}

// ------------------------------------------------------------------------------------------------
// matching
// ------------------------------------------------------------------------------------------------
namespace {

struct ChunkPlan {
    int stride;        // keys per array
    int col_cap;       // smem threshold entries
    size_t chunk_pairs;
};

bool use_tc(const esfm_ctx* ctx, const esfm_bank* b);

ChunkPlan plan_chunks(const esfm_bank* b, int64_t n_pairs) {
    ChunkPlan pl;
    const int padded = ((b->max_rows + kTile - 1) / kTile) * kTile;
    pl.stride = std::max(padded, kTile);
    pl.col_cap = pl.stride;
    const size_t key_bytes_per_pair = (size_t)4 * pl.stride * sizeof(u64);
    const size_t arena_bytes_per_pair = (size_t)std::max(b->max_rows, 1) * sizeof(esfm_dmatch_t);
    const size_t budget_keys = (size_t)4 << 30, budget_arena = (size_t)2 << 30;
    size_t c = std::min(budget_keys / key_bytes_per_pair, budget_arena / arena_bytes_per_pair);
    c = std::max<size_t>(1, std::min<size_t>(c, 65536));
    pl.chunk_pairs = (size_t)std::min<int64_t>((int64_t)c, std::max<int64_t>(n_pairs, 1));
    return pl;
}

int ensure_scratch(esfm_ctx* ctx, const esfm_bank* b, const ChunkPlan& pl) {
    size_t key_elems = pl.chunk_pairs * 4 * (size_t)pl.stride;
    size_t cap = ctx->keys_bytes / sizeof(u64);
    if (int rc = grow(&ctx->keys, &cap, key_elems)) return rc;
    ctx->keys_bytes = cap * sizeof(u64);
    if (b->kind == ESFM_KIND_F32X64 || use_tc(ctx, b))
        if (int rc = grow(&ctx->col_thr, &ctx->col_thr_elems, pl.chunk_pairs * (size_t)pl.stride)) return rc;
    size_t arena_need = pl.chunk_pairs * (size_t)std::max(b->max_rows, 1);
    if (int rc = grow(&ctx->arena, &ctx->arena_cap, arena_need)) return rc;
    if (ctx->pairs_cap < pl.chunk_pairs) {
        size_t c1 = ctx->pairs_cap, c2 = ctx->pairs_cap, c3 = ctx->pairs_cap;
        if (int rc = grow(&ctx->d_pairs, &c1, pl.chunk_pairs)) return rc;
        if (int rc = grow(&ctx->d_pair_off, &c2, pl.chunk_pairs)) return rc;
        if (int rc = grow(&ctx->d_pair_cnt, &c3, pl.chunk_pairs)) return rc;
        ctx->pairs_cap = pl.chunk_pairs;
    }
    return ESFM_OK;
}

// Tensor-core operand images of a bank (tc_layout.cuh: 3xTF32 hi/lo images for F32X64, FP8 +-1 images for B256), built the
// first time a tensor-core engine sweeps it.
int ensure_tc_layout(esfm_ctx* ctx, esfm_bank* b, int z_mode) {
    if (b->d_tc && b->tc_z == z_mode) return ESFM_OK;
    if (b->d_tc) {      // built for the other ORB encoding: rebuild (stream-ordered, earlier sweeps have been enqueued before)
        cudaFreeAsync(b->d_tc, ctx->stream);
        b->d_tc = nullptr;
    }
    b->tc_z = z_mode;
    const int n_tiles = b->tile_off[b->n_frames];
    b->tc_bytes = ((size_t)n_tiles + 1) * (b->kind == ESFM_KIND_F32X64 ? (size_t)kTcTileBytes : (size_t)kTc8TileBytes);
    cudaError_t e = cudaMallocAsync((void**)&b->d_tc, b->tc_bytes, ctx->stream);
    if (e != cudaSuccess) {
        const size_t want = b->tc_bytes;
        b->d_tc = nullptr;
        b->tc_bytes = 0;
        return fail(ESFM_ERR_NOMEM, "cudaMallocAsync(%zu) for the tensor-core bank failed: %s", want, cudaGetErrorString(e));
    }
    e = b->kind == ESFM_KIND_F32X64
            ? launch_pack_tc((const float*)b->d_rows, b->d_frame_rows, b->d_row_off, b->d_tile_off, b->n_frames, n_tiles, b->d_tc, ctx->stream)
            : launch_pack_tc8((const uint32_t*)b->d_rows, b->d_frame_rows, b->d_row_off, b->d_tile_off, b->n_frames, n_tiles, b->d_tc, z_mode, ctx->stream);
    if (e != cudaSuccess) return fail(ESFM_ERR_CUDA, "tensor-core pack kernel launch failed: %s", cudaGetErrorString(e));
    if (n_tiles > 0) ctx->stats.kernel_launches += 1;
    return ESFM_OK;
}

bool use_tc(const esfm_ctx* ctx, const esfm_bank* b) {
    return b->kind == ESFM_KIND_F32X64 ? ctx->l2_engine == ESFM_L2_ENGINE_TC : ctx->hamming_engine == ESFM_HAMMING_ENGINE_TC;
}

int units_per_pair(const esfm_ctx* ctx, const esfm_bank* b, size_t n_chunk_pairs) {
    if (n_chunk_pairs >= (size_t)4 * ctx->sm_count) return 1;
    const int qblock_rows = use_tc(ctx, b) ? kTile : (b->kind == ESFM_KIND_F32X64 ? kQTiles * kTile : kConsumerThreads * kHamRQ);
    const int max_blocks = std::max(1, (b->max_rows + qblock_rows - 1) / qblock_rows);
    const int want = (int)((4 * (size_t)ctx->sm_count + n_chunk_pairs - 1) / std::max<size_t>(n_chunk_pairs, 1));
    return std::max(1, std::min(want, max_blocks));
}

// Runs sweep + finalize for one chunk that is already described in ctx->d_pairs.
int run_chunk(esfm_ctx* ctx, esfm_bank* b, const ChunkPlan& pl, size_t n, double ratio, int cross_check, int32_t* knn_idx,
              float* knn_dist) {
    CUDA_TRY(cudaMemsetAsync(ctx->keys, 0xFF, n * 4 * (size_t)pl.stride * sizeof(u64), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_cursor, 0, 2 * sizeof(unsigned long long), ctx->stream));
    const bool tc = use_tc(ctx, b);
    // ORB "Z" encoding: the column index rides in the key, so every frame must have at most 2^15 rows; else the +-1 encoding
    const int zmode = (tc && b->kind == ESFM_KIND_B256 && ctx->orb_z && b->max_rows <= kTcZMaxRows) ? 1 : 0;
    if (tc) if (int rc = ensure_tc_layout(ctx, b, zmode)) return rc;
    if (b->kind == ESFM_KIND_F32X64 || tc)  // column thresholds start at "no bound yet": 0x7f7f7f7f = 3.39e38f (FFMA engine),
                                      // 0x6f6f6f6f = 7.4e28f (TC engine: below its 1e30 pad-row norm)
        // (TC engine: 0x6f6f6f6f = 7.4e28f for SURF, 0x47474747 = 51015f for ORB -- above every real value, below the pad rows)
        // (ORB "Z" encoding: 0x4b4b4b4b = 1.33e7f, above every key)
        CUDA_TRY(cudaMemsetAsync(ctx->col_thr, tc ? (b->kind == ESFM_KIND_F32X64 ? 0x6F : (zmode ? 0x4B : 0x47)) : 0x7F,
                                 n * (size_t)pl.stride * sizeof(uint32_t), ctx->stream));
    SweepParams sp{};
    sp.kmajor = b->d_kmajor;
    sp.rows_f32 = (const float*)b->d_rows;
    sp.tc_main = b->d_tc;
    sp.rows_b256 = (const uint4*)b->d_rows;
    sp.frame_rows = b->d_frame_rows;
    sp.frame_row_off = b->d_row_off;
    sp.frame_tile_off = b->d_tile_off;
    sp.pairs = ctx->d_pairs;
    sp.n_pairs = (int)n;
    sp.units_per_pair = units_per_pair(ctx, b, n);
    sp.keys = ctx->keys;
    sp.col_thr = ctx->col_thr;
    sp.stride = pl.stride;
    sp.col_cap = pl.col_cap;
    if (const char* dbg = getenv("ESFM_TC_DEBUG")) sp.debug_flags = atoi(dbg);
    sp.tc_qtiles = b->kind == ESFM_KIND_B256 ? ctx->tc_qtiles_orb : ctx->tc_qtiles;
    sp.tc_kind = zmode ? kTcKindB256Z : ((tc && b->kind == ESFM_KIND_F32X64 && ctx->surf_bf && ctx->tc_qtiles == 1) ? kTcKindF32BF : b->kind);
    if (ctx->profiling) CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    cudaError_t e = tc ? launch_sweep_l2_tc(sp, ctx->sm_count, ctx->stream)
                       : (b->kind == ESFM_KIND_F32X64 ? launch_sweep_l2(sp, ctx->sm_count, ctx->stream)
                                                      : launch_sweep_hamming(sp, ctx->sm_count, ctx->stream));
    if (e != cudaSuccess) return fail(ESFM_ERR_CUDA, "sweep kernel launch failed: %s", cudaGetErrorString(e));
    if (ctx->profiling) CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    FinalizeParams fp{};
    fp.kind = b->kind;
    fp.rows_f32 = (const float*)b->d_rows;
    fp.rows_b256 = (const uint4*)b->d_rows;
    fp.frame_rows = b->d_frame_rows;
    fp.frame_row_off = b->d_row_off;
    fp.pairs = ctx->d_pairs;
    fp.n_pairs = (int)n;
    fp.keys = ctx->keys;
    fp.stride = pl.stride;
    fp.ratio = ratio;
    fp.cross_check = cross_check ? 1 : 0;
    fp.b256_float_keys = (tc && b->kind == ESFM_KIND_B256) ? (zmode ? 2 : 1) : 0;
    fp.arena = ctx->arena;
    fp.arena_cap = ctx->arena_cap;
    fp.cursor = ctx->d_cursor;
    fp.overflow = (int*)(ctx->d_cursor + 1);
    fp.pair_off = ctx->d_pair_off;
    fp.pair_cnt = ctx->d_pair_cnt;
    fp.knn_idx = knn_idx;
    fp.knn_dist = knn_dist;
    e = launch_finalize(fp, ctx->stream);
    if (e != cudaSuccess) return fail(ESFM_ERR_CUDA, "finalize kernel launch failed: %s", cudaGetErrorString(e));
    if (ctx->profiling) CUDA_TRY(cudaEventRecord(ctx->ev[2], ctx->stream));
    ctx->stats.kernel_launches += 2;
    ctx->stats.sweep_launches += 1;
    ctx->arena_generation += 1;
    return ESFM_OK;
}

int collect_timing(esfm_ctx* ctx) {
    if (!ctx->profiling) return ESFM_OK;
    float a = 0.f, c = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(&c, ctx->ev[1], ctx->ev[2]));
    ctx->stats.last_sweep_ms = a;
    ctx->stats.last_finalize_ms = c;
    ctx->stats.sweep_ms_total += a;
    return ESFM_OK;
}

int match_pairs_impl(esfm_bank* b, const esfm_pair_t* pairs, int64_t n_pairs, double ratio, int cross_check, bool fetch,
                     esfm_results** out) {
    if (!b || !out) return fail(ESFM_ERR_INVALID, "esfm_match_pairs: NULL argument");
    *out = nullptr;
    if (!b->committed) return fail(ESFM_ERR_STATE, "esfm_match_pairs: bank not committed");
    if (n_pairs < 0 || (n_pairs > 0 && !pairs)) return fail(ESFM_ERR_INVALID, "esfm_match_pairs: bad pair list");
    if (!(ratio == ratio)) return fail(ESFM_ERR_INVALID, "ratio is NaN");
    esfm_ctx* ctx = b->ctx;
    if (int rc = set_device(ctx)) return rc;
    for (int64_t k = 0; k < n_pairs; ++k)
        if (pairs[k].query < 0 || pairs[k].query >= b->n_frames || pairs[k].train < 0 || pairs[k].train >= b->n_frames)
            return fail(ESFM_ERR_INVALID, "pair %lld = (%d,%d) out of range [0,%d)", (long long)k, pairs[k].query, pairs[k].train, b->n_frames);

    esfm_results* res = new (std::nothrow) esfm_results();
    if (!res) return fail(ESFM_ERR_NOMEM, "out of host memory");
    res->ctx = ctx;
    res->kind = b->kind;
    res->ratio = ratio;
    res->cross_check = cross_check ? 1 : 0;
    res->pairs.resize((size_t)n_pairs);
    res->counts.assign((size_t)n_pairs, 0);
    res->offsets.assign((size_t)n_pairs, 0);
    res->fetched = fetch;
    for (int64_t k = 0; k < n_pairs; ++k) {
        res->pairs[(size_t)k].q_frame = pairs[k].query;
        res->pairs[(size_t)k].t_frame = pairs[k].train;
    }
    if (n_pairs == 0) { *out = res; return ESFM_OK; }

    const ChunkPlan pl = plan_chunks(b, n_pairs);
    if (int rc = ensure_scratch(ctx, b, pl)) { esfm_results_destroy(res); return rc; }
    const size_t n_chunks = ((size_t)n_pairs + pl.chunk_pairs - 1) / pl.chunk_pairs;
    std::vector<uint32_t> order;
    std::vector<PairDesc> sorted;
    for (size_t c0 = 0; c0 < (size_t)n_pairs; c0 += pl.chunk_pairs) {
        const size_t n = std::min(pl.chunk_pairs, (size_t)n_pairs - c0);
        // Launch order inside the chunk: by TRAIN frame.  The sweep re-streams the train frame once per query block
        // (32x per pair at 8k rows), the query frame only once; CTAs that run concurrently take consecutive units, so
        // sorting by train frame makes them stream the SAME tiles and L2 serves the re-reads (the reference's loop order
        // shares the query frame instead and cost 6-15x the algorithmic DRAM traffic, profiles/traffic_bench_r1.csv).
        order.resize(n);
        for (size_t k = 0; k < n; ++k) order[k] = (uint32_t)k;
        const PairDesc* pp = res->pairs.data() + c0;
        std::stable_sort(order.begin(), order.end(), [pp](uint32_t a, uint32_t b2) {
            return pp[a].t_frame != pp[b2].t_frame ? pp[a].t_frame < pp[b2].t_frame : pp[a].q_frame < pp[b2].q_frame;
        });
        sorted.resize(n);
        for (size_t k = 0; k < n; ++k) sorted[k] = pp[order[k]];
        CUDA_TRY(cudaMemcpyAsync(ctx->d_pairs, sorted.data(), n * sizeof(PairDesc), cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.h2d_bytes += n * sizeof(PairDesc);
        if (int rc = run_chunk(ctx, b, pl, n, ratio, cross_check, nullptr, nullptr)) { esfm_results_destroy(res); return rc; }
        // counts + offsets + cursor back
        const size_t meta = n * (sizeof(int32_t) + sizeof(unsigned long long)) + 2 * sizeof(unsigned long long);
        if (int rc = grow_stage(ctx, meta)) { esfm_results_destroy(res); return rc; }
        unsigned long long* h_cur = (unsigned long long*)ctx->h_stage;
        unsigned long long* h_off = h_cur + 2;
        int32_t* h_cnt = (int32_t*)(h_off + n);
        CUDA_TRY(cudaMemcpyAsync(h_cur, ctx->d_cursor, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(h_off, ctx->d_pair_off, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(h_cnt, ctx->d_pair_cnt, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->stats.d2h_bytes += meta;
        if (int rc = collect_timing(ctx)) { esfm_results_destroy(res); return rc; }
        const unsigned long long n_matches = h_cur[0];
        if ((int)h_cur[1] != 0 || n_matches > ctx->arena_cap) { esfm_results_destroy(res); return fail(ESFM_ERR_CAPACITY, "match arena overflow (internal sizing error)"); }
        const uint64_t seg = (uint64_t)res->segments.size();
        for (size_t k = 0; k < n; ++k) {
            res->counts[c0 + order[k]] = h_cnt[k];
            res->offsets[c0 + order[k]] = (seg << 40) | (uint64_t)h_off[k];
            const PairDesc& pd = sorted[k];
            ctx->stats.comparisons += (uint64_t)b->rows[pd.q_frame] * (uint64_t)b->rows[pd.t_frame];
        }
        ctx->stats.pairs += n;
        res->total_matches += (int64_t)n_matches;
        if (fetch) {
            esfm_results::Segment sg{nullptr, (size_t)n_matches};
            if (n_matches > 0) {
                const size_t bytes = (size_t)n_matches * sizeof(esfm_dmatch_t);
                sg.ptr = (esfm_dmatch_t*)pool_acquire(ctx, bytes, nullptr);
                if (!sg.ptr) { esfm_results_destroy(res); return fail(ESFM_ERR_NOMEM, "pinned buffer of %zu bytes for the matches failed", bytes); }
                res->segments.push_back(sg);
                CUDA_TRY(cudaMemcpyAsync(sg.ptr, ctx->arena, bytes, cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                ctx->stats.d2h_bytes += bytes;
            } else {
                res->segments.push_back(sg);
            }
        } else {
            res->device_matches = n_matches;
            res->arena_generation = ctx->arena_generation;
            if (n_chunks > 1) res->arena_generation = 0;  // cannot be fetched later: arena is reused per chunk
        }
    }
    *out = res;
    return ESFM_OK;
}

void build_index(esfm_results* r) {
    if (!r->index.empty() || r->pairs.empty()) return;
    r->index.reserve(r->pairs.size() * 2);
    for (size_t k = 0; k < r->pairs.size(); ++k) {
        const uint64_t key = ((uint64_t)(uint32_t)r->pairs[k].q_frame << 32) | (uint32_t)r->pairs[k].t_frame;
        r->index.emplace(key, (int64_t)k);
    }
}

}  // namespace

extern "C" int esfm_match_pairs(esfm_bank_t* b, const esfm_pair_t* pairs, int64_t n_pairs, double ratio, int cross_check,
                                esfm_results_t** results) {
    return match_pairs_impl(b, pairs, n_pairs, ratio, cross_check, true, results);
}

extern "C" int esfm_match_pairs_device(esfm_bank_t* b, const esfm_pair_t* pairs, int64_t n_pairs, double ratio,
                                       int cross_check, esfm_results_t** results) {
    return match_pairs_impl(b, pairs, n_pairs, ratio, cross_check, false, results);
}

extern "C" int esfm_results_fetch(esfm_results_t* r) {
    if (!r) return fail(ESFM_ERR_INVALID, "results is NULL");
    if (r->fetched) return ESFM_OK;
    esfm_ctx* ctx = r->ctx;
    if (r->arena_generation == 0 || r->arena_generation != ctx->arena_generation)
        return fail(ESFM_ERR_STATE, "device-resident matches are gone (the arena was reused by a later batch or the batch spanned several chunks)");
    if (int rc = set_device(ctx)) return rc;
    esfm_results::Segment sg{nullptr, (size_t)r->device_matches};
    if (r->device_matches > 0) {
        const size_t bytes = (size_t)r->device_matches * sizeof(esfm_dmatch_t);
        sg.ptr = (esfm_dmatch_t*)pool_acquire(ctx, bytes, nullptr);
        if (!sg.ptr) return fail(ESFM_ERR_NOMEM, "pinned buffer of %zu bytes for the matches failed", bytes);
        CUDA_TRY(cudaMemcpyAsync(sg.ptr, ctx->arena, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->stats.d2h_bytes += bytes;
    }
    r->segments.assign(1, sg);
    r->fetched = true;
    return ESFM_OK;
}

extern "C" int esfm_match_all_pairs(esfm_bank_t* b, double ratio, int cross_check, esfm_results_t** results) {
    if (!b || !results) return fail(ESFM_ERR_INVALID, "esfm_match_all_pairs: NULL argument");
    std::vector<esfm_pair_t> pairs;
    const int n = b->n_frames;
    pairs.reserve((size_t)n * (size_t)(n > 0 ? n - 1 : 0) / 2);
    for (int i = 0; i < n; ++i)           // cpp_code/test/sfm.cpp:140
        for (int j = 0; j < i; ++j) {     // cpp_code/test/sfm.cpp:143
            esfm_pair_t p;
            p.query = i;                  // frames[i] is cur_frame_1 = query  (sfm.cpp:153,156)
            p.train = j;
            pairs.push_back(p);
        }
    return match_pairs_impl(b, pairs.data(), (int64_t)pairs.size(), ratio, cross_check, true, results);
}

extern "C" int esfm_match_pair(esfm_bank_t* b, int query_frame, int train_frame, double ratio, int cross_check,
                               esfm_dmatch_t* out, int cap, int* n_matches) {
    if (!n_matches) return fail(ESFM_ERR_INVALID, "n_matches is NULL");
    *n_matches = 0;
    esfm_pair_t p;
    p.query = query_frame;
    p.train = train_frame;
    esfm_results* r = nullptr;
    if (int rc = match_pairs_impl(b, &p, 1, ratio, cross_check, true, &r)) return rc;
    const int n = r->counts[0];
    if (n > cap || (n > 0 && !out)) {
        esfm_results_destroy(r);
        return fail(ESFM_ERR_CAPACITY, "output buffer holds %d matches, %d needed", cap, n);
    }
    if (n > 0) memcpy(out, r->segments[0].ptr + (r->offsets[0] & (((uint64_t)1 << 40) - 1)), (size_t)n * sizeof(esfm_dmatch_t));
    *n_matches = n;
    esfm_results_destroy(r);
    return ESFM_OK;
}

extern "C" int esfm_match_descriptors(esfm_ctx_t* ctx, esfm_kind kind, const void* query, int rows_q, size_t step_q,
                                      const void* train, int rows_t, size_t step_t, int cols, double ratio, int cross_check,
                                      esfm_dmatch_t* out, int cap, int* n_matches) {
    if (!n_matches) return fail(ESFM_ERR_INVALID, "n_matches is NULL");
    *n_matches = 0;
    esfm_bank* b = nullptr;
    if (int rc = esfm_bank_create(ctx, kind, 2, &b)) return rc;
    int rc = esfm_bank_set_frame(b, 0, query, rows_q, cols, step_q);
    if (!rc) rc = esfm_bank_set_frame(b, 1, train, rows_t, cols, step_t);
    if (!rc) rc = esfm_bank_commit(b);
    if (!rc) rc = esfm_match_pair(b, 0, 1, ratio, cross_check, out, cap, n_matches);
    esfm_bank_destroy(b);
    return rc;
}

extern "C" int esfm_knn2_pair(esfm_bank_t* b, int query_frame, int train_frame, int32_t* idx, float* dist) {
    if (!b || !idx || !dist) return fail(ESFM_ERR_INVALID, "esfm_knn2_pair: NULL argument");
    if (!b->committed) return fail(ESFM_ERR_STATE, "bank not committed");
    if (query_frame < 0 || query_frame >= b->n_frames || train_frame < 0 || train_frame >= b->n_frames)
        return fail(ESFM_ERR_INVALID, "frame index out of range");
    esfm_ctx* ctx = b->ctx;
    if (int rc = set_device(ctx)) return rc;
    const int fq = b->rows[query_frame];
    if (fq == 0) return ESFM_OK;
    const ChunkPlan pl = plan_chunks(b, 1);
    if (int rc = ensure_scratch(ctx, b, pl)) return rc;
    int32_t* d_idx = nullptr;
    float* d_dist = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_idx, (size_t)fq * 2 * sizeof(int32_t)));
    cudaError_t e = cudaMalloc((void**)&d_dist, (size_t)fq * 2 * sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_idx); return fail(ESFM_ERR_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    PairDesc pd;
    pd.q_frame = query_frame;
    pd.t_frame = train_frame;
    int rc = ESFM_OK;
    e = cudaMemcpyAsync(ctx->d_pairs, &pd, sizeof pd, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) rc = fail(ESFM_ERR_CUDA, "cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
    if (!rc) rc = run_chunk(ctx, b, pl, 1, 0.0, 0, d_idx, d_dist);
    if (!rc) {
        e = cudaMemcpyAsync(idx, d_idx, (size_t)fq * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dist, d_dist, (size_t)fq * 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ESFM_ERR_CUDA, "knn2 copy back failed: %s", cudaGetErrorString(e));
        ctx->stats.d2h_bytes += (size_t)fq * 16;
        ctx->stats.pairs += 1;
        ctx->stats.comparisons += (uint64_t)fq * (uint64_t)b->rows[train_frame];
    }
    cudaFree(d_idx);
    cudaFree(d_dist);
    if (!rc) rc = collect_timing(ctx);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------------
extern "C" int esfm_results_counts(esfm_results_t* r, int64_t* n_pairs, int64_t* n_matches) {
    if (!r) return fail(ESFM_ERR_INVALID, "results is NULL");
    if (n_pairs) *n_pairs = (int64_t)r->pairs.size();
    if (n_matches) *n_matches = r->total_matches;
    return ESFM_OK;
}

extern "C" int esfm_results_pair_at(esfm_results_t* r, int64_t k, int* query_frame, int* train_frame,
                                    const esfm_dmatch_t** matches, int* n_matches) {
    if (!r) return fail(ESFM_ERR_INVALID, "results is NULL");
    if (k < 0 || k >= (int64_t)r->pairs.size()) return fail(ESFM_ERR_INVALID, "pair index %lld out of range", (long long)k);
    if (query_frame) *query_frame = r->pairs[(size_t)k].q_frame;
    if (train_frame) *train_frame = r->pairs[(size_t)k].t_frame;
    if (n_matches) *n_matches = r->counts[(size_t)k];
    if (matches) {
        if (!r->fetched) return fail(ESFM_ERR_STATE, "matches are device-resident; call esfm_results_fetch first");
        const uint64_t o = r->offsets[(size_t)k];
        *matches = r->counts[(size_t)k] > 0 ? r->segments[(size_t)(o >> 40)].ptr + (o & (((uint64_t)1 << 40) - 1)) : nullptr;
    }
    return ESFM_OK;
}

extern "C" int esfm_results_pair(esfm_results_t* r, int query_frame, int train_frame, const esfm_dmatch_t** matches,
                                 int* n_matches) {
    if (!r) return fail(ESFM_ERR_INVALID, "results is NULL");
    build_index(r);
    const uint64_t key = ((uint64_t)(uint32_t)query_frame << 32) | (uint32_t)train_frame;
    auto it = r->index.find(key);
    if (it == r->index.end()) return fail(ESFM_ERR_INVALID, "pair (%d,%d) is not part of this batch", query_frame, train_frame);
    return esfm_results_pair_at(r, it->second, nullptr, nullptr, matches, n_matches);
}

extern "C" int esfm_results_pair_counts(esfm_results_t* r, int32_t* counts) {
    if (!r || !counts) return fail(ESFM_ERR_INVALID, "NULL argument");
    if (!r->counts.empty()) memcpy(counts, r->counts.data(), r->counts.size() * sizeof(int32_t));
    return ESFM_OK;
}

extern "C" int esfm_results_destroy(esfm_results_t* r) {
    if (!r) return ESFM_OK;
    for (auto& s : r->segments) {
        if (r->heap_segments) free(s.ptr);
        else pool_release(r->ctx, s.ptr);
    }
    delete r;
    return ESFM_OK;
}

// ------------------------------------------------------------------------------------------------
// persistence of a fetched batch (host only)
// ------------------------------------------------------------------------------------------------
namespace {
struct MatchFileHeader {
    char magic[8];
    uint32_t version, dmatch_bytes;
    int64_t n_pairs, n_matches;
    int32_t kind, cross_check;
    double ratio;
};
static_assert(sizeof(MatchFileHeader) == 48, "match file header layout");
const char kMatchMagic[8] = {'E', 'S', 'F', 'M', 'M', 'T', 'C', 'H'};
}  // namespace

extern "C" int esfm_results_save(esfm_results_t* r, const char* path) {
    if (!r || !path) return fail(ESFM_ERR_INVALID, "esfm_results_save: NULL argument");
    if (!r->fetched) return fail(ESFM_ERR_STATE, "matches are device-resident; call esfm_results_fetch first");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(ESFM_ERR_INVALID, "esfm_results_save: cannot open %s for writing", path);
    MatchFileHeader h{};
    memcpy(h.magic, kMatchMagic, 8);
    h.version = 1;
    h.dmatch_bytes = (uint32_t)sizeof(esfm_dmatch_t);
    h.n_pairs = (int64_t)r->pairs.size();
    h.n_matches = 0;
    for (int32_t c : r->counts) h.n_matches += c;
    h.kind = r->kind;
    h.cross_check = r->cross_check;
    h.ratio = r->ratio;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    static_assert(sizeof(PairDesc) == sizeof(esfm_pair_t), "pair layout");
    if (ok && h.n_pairs) ok = fwrite(r->pairs.data(), sizeof(PairDesc), r->pairs.size(), f) == r->pairs.size();
    if (ok && h.n_pairs) ok = fwrite(r->counts.data(), sizeof(int32_t), r->counts.size(), f) == r->counts.size();
    for (size_t k = 0; ok && k < r->pairs.size(); ++k) {
        const int32_t n = r->counts[k];
        if (n <= 0) continue;
        const uint64_t o = r->offsets[k];
        const esfm_dmatch_t* m = r->segments[(size_t)(o >> 40)].ptr + (o & (((uint64_t)1 << 40) - 1));
        ok = fwrite(m, sizeof(esfm_dmatch_t), (size_t)n, f) == (size_t)n;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(ESFM_ERR_INVALID, "esfm_results_save: short write to %s", path);
    return ESFM_OK;
}

extern "C" int esfm_results_params(esfm_results_t* r, int* kind, double* ratio, int* cross_check) {
    if (!r) return fail(ESFM_ERR_INVALID, "results is NULL");
    if (kind) *kind = r->kind;
    if (ratio) *ratio = r->ratio;
    if (cross_check) *cross_check = r->cross_check;
    return ESFM_OK;
}

extern "C" int esfm_results_load(const char* path, esfm_results_t** out) {
    if (!path || !out) return fail(ESFM_ERR_INVALID, "esfm_results_load: NULL argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(ESFM_ERR_INVALID, "esfm_results_load: cannot open %s", path);
    MatchFileHeader h{};
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kMatchMagic, 8) != 0 || h.version != 1 ||
        h.dmatch_bytes != sizeof(esfm_dmatch_t) || h.n_pairs < 0 || h.n_matches < 0) {
        fclose(f);
        return fail(ESFM_ERR_INVALID, "esfm_results_load: %s is not an esfm match file (version 1)", path);
    }
    esfm_results* r = new (std::nothrow) esfm_results();
    if (!r) { fclose(f); return fail(ESFM_ERR_NOMEM, "out of host memory"); }
    r->heap_segments = true;
    r->fetched = true;
    r->kind = h.kind;
    r->cross_check = h.cross_check;
    r->ratio = h.ratio;
    bool ok = true;
    try {
        r->pairs.resize((size_t)h.n_pairs);
        r->counts.resize((size_t)h.n_pairs);
        r->offsets.resize((size_t)h.n_pairs);
    } catch (...) { ok = false; }
    if (ok && h.n_pairs) ok = fread(r->pairs.data(), sizeof(PairDesc), r->pairs.size(), f) == r->pairs.size();
    if (ok && h.n_pairs) ok = fread(r->counts.data(), sizeof(int32_t), r->counts.size(), f) == r->counts.size();
    int64_t total = 0;
    for (size_t k = 0; ok && k < r->pairs.size(); ++k) {
        if (r->counts[k] < 0) { ok = false; break; }
        r->offsets[k] = (uint64_t)total;      // segment 0
        total += r->counts[k];
    }
    if (ok && total != h.n_matches) ok = false;
    if (ok) {
        esfm_results::Segment sg{nullptr, (size_t)total};
        if (total > 0) {
            sg.ptr = (esfm_dmatch_t*)malloc((size_t)total * sizeof(esfm_dmatch_t));
            ok = sg.ptr && fread(sg.ptr, sizeof(esfm_dmatch_t), (size_t)total, f) == (size_t)total;
        }
        r->segments.push_back(sg);
        r->total_matches = total;
    }
    fclose(f);
    if (!ok) {
        esfm_results_destroy(r);
        return fail(ESFM_ERR_INVALID, "esfm_results_load: %s is truncated or inconsistent", path);
    }
    *out = r;
    return ESFM_OK;
}
