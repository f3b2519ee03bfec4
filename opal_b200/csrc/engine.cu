// engine.cu -- host-side engine: database packing, geometry choice, precision ladder, launches.
//
// Plays the role of the reference's escalation drivers searchDatabaseSW / searchDatabase<MODE>
// (reference src/opal.cpp:496-535, 983-1021) and of lane refill (loadNextSequence, :472-490), but
// scheduled for a GPU: the database is sorted by length once (longest first), adjacent targets
// are paired into the two 16-bit lanes of a group, warps pull groups of similar length from a
// global counter, and the 16 -> 32 bit escalation re-runs exactly the flagged targets.
#include "engine.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "r_list.h"
#include "search_kernel.cuh"

namespace opalb200 {

// ------------------------------------------------------------------ errors
static thread_local std::string g_lastError;
void set_error(const std::string& msg) { g_lastError = msg; }
const char* last_error() { return g_lastError.c_str(); }

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
            return false;                                                                    \
        }                                                                                    \
    } while (0)

// ------------------------------------------------------------------ kernel registry
#define OPAL_DECLARE_TABLE(R) const void* const* kernel_table_R##R();
OPAL_R_LIST(OPAL_DECLARE_TABLE)
#undef OPAL_DECLARE_TABLE

const std::vector<KernelTable>& kernel_tables() {
    static const std::vector<KernelTable> tables = {
#define OPAL_TABLE_ENTRY(R) {R, kernel_table_R##R()},
        OPAL_R_LIST(OPAL_TABLE_ENTRY)
#undef OPAL_TABLE_ENTRY
    };
    return tables;
}

// ------------------------------------------------------------------ per-device resource cache
// opalSearchDatabase packs and uploads its database on every call (the reference's signature takes
// host pointers each time), so the CUDA allocations behind a call are recycled instead of being
// returned to the driver: cudaMalloc / cudaMallocHost / stream and event creation cost far more
// than the search itself on a 12k-sequence database.  Blocks are matched by size (first cached block
// within 2x of the request); the cache is bounded and thread-safe.
namespace {
struct DeviceInfo { bool known = false; int numSMs = 0, smemLimit = 0, major = 0; };
struct CachedBlock { void* p; size_t bytes; };
struct ResourceCache {
    std::mutex mu;
    std::vector<CachedBlock> freeDevice[16], freePinned;
    std::unordered_map<void*, size_t> liveBytes;
    std::vector<cudaStream_t> streams[16];
    std::vector<cudaEvent_t> events[16];
    DeviceInfo info[16];
    size_t cachedDevice[16] = {0}, cachedPinned = 0;
};
ResourceCache& cache() { static ResourceCache* c = new ResourceCache(); return *c; }  // leaked on purpose: no teardown order issues
constexpr size_t kMaxCachedDevice = 8ull << 30, kMaxCachedPinned = 2ull << 30;

bool take_block(std::vector<CachedBlock>& list, size_t bytes, void** out, size_t* got) {
    size_t best = list.size();
    for (size_t i = 0; i < list.size(); i++)
        if (list[i].bytes >= bytes && list[i].bytes <= 2 * bytes + (1 << 16) && (best == list.size() || list[i].bytes < list[best].bytes)) best = i;
    if (best == list.size()) return false;
    *out = list[best].p; *got = list[best].bytes;
    list[best] = list.back(); list.pop_back();
    return true;
}
}  // namespace

static bool device_alloc(int device, void** p, size_t bytes) {
    bytes = std::max<size_t>(bytes, 256);
    ResourceCache& c = cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        size_t got = 0;
        if (take_block(c.freeDevice[device & 15], bytes, p, &got)) { c.cachedDevice[device & 15] -= got; c.liveBytes[*p] = got; return true; }
    }
    CUDA_TRY(cudaMalloc(p, bytes));
    std::lock_guard<std::mutex> lk(c.mu);
    c.liveBytes[*p] = bytes;
    return true;
}
static void device_release(int device, void* p) {
    if (!p) return;
    ResourceCache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    const size_t bytes = c.liveBytes[p];
    c.liveBytes.erase(p);
    if (c.cachedDevice[device & 15] + bytes > kMaxCachedDevice) { cudaFree(p); return; }
    c.freeDevice[device & 15].push_back({p, bytes});
    c.cachedDevice[device & 15] += bytes;
}
static bool pinned_alloc(void** p, size_t bytes) {
    bytes = std::max<size_t>(bytes, 256);
    ResourceCache& c = cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        size_t got = 0;
        if (take_block(c.freePinned, bytes, p, &got)) { c.cachedPinned -= got; c.liveBytes[*p] = got; return true; }
    }
    CUDA_TRY(cudaMallocHost(p, bytes));
    std::lock_guard<std::mutex> lk(c.mu);
    c.liveBytes[*p] = bytes;
    return true;
}
static void pinned_release(void* p) {
    if (!p) return;
    ResourceCache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    const size_t bytes = c.liveBytes[p];
    c.liveBytes.erase(p);
    if (c.cachedPinned + bytes > kMaxCachedPinned) { cudaFreeHost(p); return; }
    c.freePinned.push_back({p, bytes});
    c.cachedPinned += bytes;
}
static bool stream_acquire(int device, cudaStream_t* s) {
    ResourceCache& c = cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto& v = c.streams[device & 15];
        if (!v.empty()) { *s = v.back(); v.pop_back(); return true; }
    }
    CUDA_TRY(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking));
    return true;
}
static void stream_release(int device, cudaStream_t s) {
    if (!s) return;
    ResourceCache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    c.streams[device & 15].push_back(s);
}
static bool event_acquire(int device, cudaEvent_t* e) {
    ResourceCache& c = cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto& v = c.events[device & 15];
        if (!v.empty()) { *e = v.back(); v.pop_back(); return true; }
    }
    CUDA_TRY(cudaEventCreate(e));
    return true;
}
static void event_release(int device, cudaEvent_t e) {
    if (!e) return;
    ResourceCache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    c.events[device & 15].push_back(e);
}
static bool device_info(int device, DeviceInfo* out) {
    ResourceCache& c = cache();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        if (c.info[device & 15].known) { *out = c.info[device & 15]; return true; }
    }
    DeviceInfo di;
    CUDA_TRY(cudaDeviceGetAttribute(&di.numSMs, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&di.smemLimit, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_TRY(cudaDeviceGetAttribute(&di.major, cudaDevAttrComputeCapabilityMajor, device));
    di.known = true;
    std::lock_guard<std::mutex> lk(c.mu);
    c.info[device & 15] = di;
    *out = di;
    return true;
}

// Thread stride in the profile: a multiple of 4 words (16-byte aligned LDS.128) with an odd number of
// 16-byte units, so the 8 lanes of a quarter-warp hit 8 different bank groups whatever their residues are.
static int rpad_of(int R) { const int r4 = (R + 3) / 4; return 4 * ((r4 % 2 == 0) ? r4 + 1 : r4); }

// Picks (G, R, passes, warps per scheduler partition) for a query of Q rows.
//
// Timing model (cycles), calibrated on B200 runs (profiles/): the integer pipe retires one packed DPX
// instruction per 2 cycles per partition and a cell pair costs ~5.5 of them (11 cycles per row ideal,
// 14.5 measured with the profile loads and adds around them); a step of one warp alone costs another
// ~160 cycles of exposed latency (exchange, residue fetch, profile load).  With k warps sharing a
// partition that latency is hidden, so one step of one warp lasts max(k * 14.5 R, 14.5 R + 160).  Two bounds:
//   throughput:  sum over warp-tasks of steps * stepTime / (partitions * k)
//   tail:        the longest target's steps * stepTime  (it cannot be split across warps)
// Small databases with a long tail (BASELINE configs[1]) are tail-bound and want G = 32 and k = 1;
// large ones are throughput-bound and want few threads per target and k = 4.
static bool pick_geometry(int Q, int A, int lanes, const std::vector<int>& lens, int smemLimit, int numSMs, int mode,
                          Geometry* out) {
    const int planes = lanes == 2 ? 2 : 1;
    const auto& tables = kernel_tables();
    double bestCost = 1e300;
    bool found = false;
    // sum of lengths and the longest length; lens is sorted longest first and tasks pair neighbours
    double sumLen = 0;
    for (size_t i = 0; i < lens.size(); i += lanes) sumLen += lens[i];
    const double tasks = (double)((lens.size() + lanes - 1) / lanes);
    const double maxLen = lens.empty() ? 0 : lens[0];
    for (size_t ti = 0; ti < tables.size(); ti++) {
        const int R = tables[ti].R;
        const int Rpad = rpad_of(R);
        for (int G = 1; G <= 32; G *= 2) {
            const int rowStride = (G * Rpad + 31) / 32 * 32;
            const size_t smem = (size_t)planes * (A + 1) * rowStride * 4;
            if (smem > (size_t)smemLimit) continue;
            const int rows = G * R;
            const int passes = (Q + rows - 1) / rows;
            const double groupsPerWarp = 32.0 / G;
            const double warpSteps = (sumLen + tasks * (G - 1)) / groupsPerWarp;  // per pass
            const double warpTasks = std::max(1.0, tasks / groupsPerWarp);
            for (int k = 1; k <= kBlockThreads / 128; k *= 2) {
                // with fewer warp-tasks than resident warps the partitions are not shared k ways
                const double kEff = std::max(1.0, std::min((double)k, std::ceil(warpTasks / (numSMs * 4.0))));
                const double stepTime = std::max(kEff * 14.5 * R, 14.5 * R + 160.0);
                const double warpsBusy = std::min((double)numSMs * 4 * k, warpTasks);
                const double throughput = warpSteps * stepTime / warpsBusy;
                const double tail = (maxLen + G - 1) * stepTime;
                const double cost = passes * (std::max(throughput, tail) + 0.15 * std::min(throughput, tail) + 4000.0);
                if (cost < bestCost) {
                    bestCost = cost;
                    found = true;
                    out->G = G; out->R = R; out->tableIndex = (int)ti; out->passes = passes; out->Rpad = Rpad;
                    out->rowStride = rowStride; out->smemBytes = smem; out->warpsPerPartition = k;
                    out->padTop = (mode == kModeNW) ? 0 : passes * rows - Q;
                }
            }
        }
    }
    // Development override: OPAL_B200_GEOMETRY="G,R,k" forces a geometry (ignored when it does not fit).
    if (const char* env = getenv("OPAL_B200_GEOMETRY")) {
        int G = 0, R = 0, k = 0;
        if (sscanf(env, "%d,%d,%d", &G, &R, &k) == 3 && G >= 1 && G <= 32 && (G & (G - 1)) == 0 && k >= 1 && k <= kBlockThreads / 128) {
            for (size_t ti = 0; ti < tables.size(); ti++) {
                if (tables[ti].R != R) continue;
                const int Rpad = rpad_of(R), rowStride = (G * Rpad + 31) / 32 * 32;
                const size_t smem = (size_t)planes * (A + 1) * rowStride * 4;
                if (smem > (size_t)smemLimit) break;
                const int rows = G * R, passes = (Q + rows - 1) / rows;
                out->G = G; out->R = R; out->tableIndex = (int)ti; out->passes = passes; out->Rpad = Rpad;
                out->rowStride = rowStride; out->smemBytes = smem; out->warpsPerPartition = k;
                out->padTop = (mode == kModeNW) ? 0 : passes * rows - Q;
                found = true;
            }
        }
    }
    if (!found) set_error("alphabet too large for the shared-memory query profile");
    return found;
}

// ---------------------------------------------------------------- database packing
// Builds the paired stream from the plain length-sorted residues: one warp per target pair.
static __global__ void pack_pairs_kernel(const uint8_t* residues, const long long* offsets, const int* lengths, int numTargets,
                                  const long long* pairOffsets, int numPairs, uint16_t* pairStream, int* maxCode) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= numPairs) return;
    const int a = 2 * warp, b = 2 * warp + 1;
    const uint8_t* sa = residues + offsets[a];
    const int Ta = lengths[a];
    const uint8_t* sb = residues;
    int Tb = 0;
    if (b < numTargets) { sb = residues + offsets[b]; Tb = lengths[b]; }
    uint16_t* out = pairStream + pairOffsets[warp];
    uint32_t mx = 0;
    for (int c = lane; c < Ta; c += 32) {
        const uint32_t lo = (uint32_t)sa[c] + 1u;
        const uint32_t hi = c < Tb ? (uint32_t)sb[c] + 1u : 0u;
        mx = max(mx, max(lo, hi));
        out[c] = (uint16_t)(lo | (hi << 8));
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0 && mx > 0) atomicMax(maxCode, (int)mx - 1);  // largest residue code seen
}

// ------------------------------------------------------------------ DeviceDb
DeviceDb* DeviceDb::create(unsigned char* const* db, int n, const int* lens, int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_error("no usable CUDA device");
        return nullptr;
    }
    DeviceDb* d = new DeviceDb();
    d->device_ = device;
    d->n_ = n;
    auto fail = [&]() -> DeviceDb* { delete d; return nullptr; };
    auto ok = [&]() -> bool {
        CUDA_TRY(cudaSetDevice(device));
        DeviceInfo di;
        if (!device_info(device, &di)) return false;
        if (di.major < 10) { set_error("device is not sm_100 or newer"); return false; }
        d->numSMs_ = di.numSMs;
        d->smemLimit_ = di.smemLimit;
        if (!stream_acquire(device, &d->stream_) || !event_acquire(device, &d->evStart_) || !event_acquire(device, &d->evStop_)) return false;

        // ---- sort by length, longest first (counting sort; stable in caller order)
        int maxLen = 0;
        for (int i = 0; i < n; i++) {
            if (lens[i] < 0) { set_error("negative sequence length"); return false; }
            maxLen = std::max(maxLen, lens[i]);
        }
        d->order_.resize(n);
        d->pos_.resize(n);
        if (maxLen <= (1 << 22)) {
            std::vector<int> start(maxLen + 2, 0);
            for (int i = 0; i < n; i++) start[maxLen - lens[i] + 1]++;
            for (int k = 1; k <= maxLen + 1; k++) start[k] += start[k - 1];
            for (int i = 0; i < n; i++) d->order_[start[maxLen - lens[i]]++] = i;
        } else {
            for (int i = 0; i < n; i++) d->order_[i] = i;
            std::stable_sort(d->order_.begin(), d->order_.end(), [&](int a, int b) { return lens[a] > lens[b]; });
        }
        d->sortedLen_.resize(n);
        d->offsets_.resize((size_t)n + 1);
        long long total = 0;
        for (int p = 0; p < n; p++) {
            const int i = d->order_[p];
            d->pos_[i] = p;
            d->sortedLen_[p] = lens[i];
            d->offsets_[p] = total;
            total += lens[i];
        }
        d->offsets_[n] = total;
        d->totalResidues_ = total;

        // ---- gather into pinned staging, in sorted order, on several host threads
        const size_t bytes = (size_t)total + 64;
        uint8_t* staging = nullptr;
        if (!pinned_alloc((void**)&staging, bytes)) return false;
        memset(staging + total, 0, 64);
        int nThreads = (int)std::min<long long>(std::max(1u, std::thread::hardware_concurrency()), 1 + total / (1 << 20));
        nThreads = std::max(1, std::min(nThreads, 32));
        auto gather = [&](int lo, int hi) {
            for (int p = lo; p < hi; p++)
                if (d->sortedLen_[p] > 0) memcpy(staging + d->offsets_[p], db[d->order_[p]], (size_t)d->sortedLen_[p]);
        };
        if (nThreads == 1) gather(0, n);
        else {
            std::vector<std::thread> th;
            int lo = 0;
            for (int k = 0; k < nThreads; k++) {
                const long long target = total * (k + 1) / nThreads;
                int hi = (k == nThreads - 1) ? n : (int)(std::upper_bound(d->offsets_.begin(), d->offsets_.begin() + n, target) - d->offsets_.begin());
                hi = std::max(hi, lo);
                th.emplace_back(gather, lo, hi);
                lo = hi;
            }
            for (auto& t : th) t.join();
        }
        if (!device_alloc(device, (void**)&d->dResidues_, bytes)) return false;
        CUDA_TRY(cudaMemcpyAsync(d->dResidues_, staging, bytes, cudaMemcpyHostToDevice, d->stream_));
        if (!device_alloc(device, (void**)&d->dOffsets_, sizeof(long long) * ((size_t)n + 1))) return false;
        CUDA_TRY(cudaMemcpyAsync(d->dOffsets_, d->offsets_.data(), sizeof(long long) * ((size_t)n + 1), cudaMemcpyHostToDevice, d->stream_));
        const size_t nInts = sizeof(int) * (size_t)std::max(n, 1);
        if (!device_alloc(device, (void**)&d->dLengths_, nInts)) return false;
        CUDA_TRY(cudaMemcpyAsync(d->dLengths_, d->sortedLen_.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, d->stream_));
        if (!device_alloc(device, (void**)&d->dScore_, nInts)) return false;
        if (!device_alloc(device, (void**)&d->dEndQ_, nInts)) return false;
        if (!device_alloc(device, (void**)&d->dEndT_, nInts)) return false;
        if (!device_alloc(device, (void**)&d->dTaskList_, nInts)) return false;
        if (!device_alloc(device, (void**)&d->dCounters_, sizeof(int) * 256)) return false;
        if (!pinned_alloc((void**)&d->hScore_, nInts)) return false;
        if (!pinned_alloc((void**)&d->hEndQ_, nInts)) return false;
        if (!pinned_alloc((void**)&d->hEndT_, nInts)) return false;
        // ---- paired stream: [32 zeros][pair 0 columns][32 zeros][pair 1 columns] ... built on the device
        d->numPairs_ = (n + 1) / 2;
        std::vector<long long> pairOff((size_t)std::max(d->numPairs_, 1));
        long long entries = 32;
        for (int p = 0; p < d->numPairs_; p++) { pairOff[p] = entries; entries += d->sortedLen_[2 * p] + 32; }
        entries += 64;
        if (!device_alloc(device, (void**)&d->dPairStream_, sizeof(uint16_t) * (size_t)entries)) return false;
        if (!device_alloc(device, (void**)&d->dPairOffsets_, sizeof(long long) * pairOff.size())) return false;
        if (!device_alloc(device, (void**)&d->dMaxCode_, sizeof(int))) return false;
        CUDA_TRY(cudaMemsetAsync(d->dPairStream_, 0, sizeof(uint16_t) * (size_t)entries, d->stream_));
        CUDA_TRY(cudaMemsetAsync(d->dMaxCode_, 0, sizeof(int), d->stream_));
        CUDA_TRY(cudaMemcpyAsync(d->dPairOffsets_, pairOff.data(), sizeof(long long) * pairOff.size(), cudaMemcpyHostToDevice, d->stream_));
        if (d->numPairs_ > 0) {
            pack_pairs_kernel<<<(d->numPairs_ + 7) / 8, 256, 0, d->stream_>>>(d->dResidues_, d->dOffsets_, d->dLengths_, n, d->dPairOffsets_,
                                                                             d->numPairs_, d->dPairStream_, d->dMaxCode_);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaMemcpyAsync(&d->maxCode_, d->dMaxCode_, sizeof(int), cudaMemcpyDeviceToHost, d->stream_));
        CUDA_TRY(cudaStreamSynchronize(d->stream_));
        pinned_release(staging);
        return true;
    }();
    return ok ? d : fail();
}

DeviceDb::~DeviceDb() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    void* dev[] = {dResidues_, dOffsets_, dLengths_, dScore_, dEndQ_, dEndT_, dTaskList_, dCounters_, dBndH_, dBndF_, dQuery_, dMatrix_, dPairStream_, dPairOffsets_, dMaxCode_};
    for (void* p : dev) device_release(device_, p);
    pinned_release(hScore_); pinned_release(hEndQ_); pinned_release(hEndT_);
    event_release(device_, evStart_); event_release(device_, evStop_);
    stream_release(device_, stream_);
}

bool DeviceDb::ensure_boundary() {
    if (dBndH_) return true;
    const size_t bytes = 2 * sizeof(uint32_t) * (size_t)(totalResidues_ + 64);  // two halves: ping-pong between passes
    if (!device_alloc(device_, (void**)&dBndH_, bytes) || !device_alloc(device_, (void**)&dBndF_, bytes)) return false;
    return true;
}

// Runs every pass of one precision class over `list` (sorted positions, longest first).
int DeviceDb::run_class(int type, const std::vector<int>& list, const unsigned char* dQuery, const int* dMatrix, int Q,
                        int Go, int Ge, int A, int wantEnd, int mode, int maxScore, int* launchSlot) {
    if (list.empty()) return 0;
    const int lanes = type == 0 ? 2 : 1;
    std::vector<int> lens(list.size());
    for (size_t k = 0; k < list.size(); k++) lens[k] = sortedLen_[list[k]];
    Geometry g;
    if (!pick_geometry(Q, A, lanes, lens, smemLimit_, numSMs_, mode, &g)) return OPAL_B200_ERR_CUDA;
    if (g.passes > 1 && !ensure_boundary()) return OPAL_B200_ERR_CUDA;
    if (*launchSlot + g.passes > 256) { set_error("too many passes"); return OPAL_B200_ERR_CUDA; }

    // Packed16 works on the pairs fixed at packing time (targets 2p, 2p+1): a pair runs if either member
    // is wanted; results of unwanted members are simply not published.
    std::vector<int> tasks;
    if (lanes == 2) {
        for (int t : list) if (tasks.empty() || tasks.back() != (t >> 1)) tasks.push_back(t >> 1);
    } else {
        tasks = list;
    }
    const bool identity = (int)tasks.size() == (lanes == 2 ? numPairs_ : n_);
    auto okc = [&]() -> bool {
        if (!identity)
            CUDA_TRY(cudaMemcpyAsync(dTaskList_, tasks.data(), sizeof(int) * tasks.size(), cudaMemcpyHostToDevice, stream_));
        // SW score+end at 16 bits uses the key-tracking flavor: exact below fastEndLimit, and warps that meet a
        // larger score sweep their tasks again with the exact per-cell tracking inside the same kernel.
        const int fastEndLimit = (32768 >> kRowBits) - std::max(maxScore, 0) - 1;
        const bool fastEnd = mode == kModeSW && wantEnd && type == 0 && g.R <= (1 << kRowBits) && fastEndLimit >= 64 &&
                             !getenv("OPAL_B200_EXACT_END");
        const int flavor = (mode == kModeSW) ? (wantEnd ? (fastEnd ? kFlavorSWEndFast : kFlavorSWEnd) : kFlavorSWScore) : kFlavorGlobal;
        const void* fn = kernel_tables()[g.tableIndex].fn[type * 4 + flavor];
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smemBytes));
        for (int pass = 0; pass < g.passes; pass++) {
            SearchParams p;
            memset(&p, 0, sizeof(p));
            p.query = dQuery; p.matrix = dMatrix; p.Q = Q; p.A = A; p.gapOpen = Go; p.gapExt = Ge; p.mode = mode;
            p.wantEnd = wantEnd;
            p.G = g.G; p.rowBase = pass * g.G * g.R; p.padTop = g.padTop; p.pass = pass; p.numPasses = g.passes;
            p.rowStride = g.rowStride; p.Rpad = g.Rpad;
            p.residues = dResidues_; p.offsets = dOffsets_; p.lengths = dLengths_;
            p.pairStream = dPairStream_; p.pairOffsets = dPairOffsets_; p.numTargets = n_;
            p.taskList = identity ? nullptr : dTaskList_;
            p.numTasks = (int)tasks.size();
            p.counter = dCounters_ + (*launchSlot)++;
            const size_t half = (size_t)(totalResidues_ + 64);
            p.bndInH = dBndH_ ? dBndH_ + ((pass + 1) & 1) * half : nullptr; p.bndInF = dBndF_ ? dBndF_ + ((pass + 1) & 1) * half : nullptr;
            p.bndOutH = dBndH_ ? dBndH_ + (pass & 1) * half : nullptr; p.bndOutF = dBndF_ ? dBndF_ + (pass & 1) * half : nullptr;
            p.outScore = dScore_; p.outEndQ = dEndQ_; p.outEndT = dEndT_;
            p.overflowLimit = type == 0 ? 32767 - std::max(maxScore, 0) - 1 : (1 << 30);
            p.padLetterScore = type == 0 ? -16384 : 0;
            p.one = 1; p.keyScale = 1 << kRowBits; p.fastEndLimit = fastEndLimit;
            void* args[] = {&p};
            const int warpsPerBlock = 4 * g.warpsPerPartition;  // one block per SM, k warps per scheduler partition
            const long long warpsNeeded = ((long long)tasks.size() * g.G + 31) / 32;
            const int blocks = (int)std::max<long long>(1, std::min<long long>(numSMs_, (warpsNeeded + warpsPerBlock - 1) / warpsPerBlock));
            CUDA_TRY(cudaLaunchKernel(fn, dim3(blocks), dim3(32 * warpsPerBlock), args, g.smemBytes, stream_));
            stats_.kernelLaunches++;
        }
        return true;
    }();
    if (!okc) return OPAL_B200_ERR_CUDA;
    stats_.G = g.G; stats_.R = g.R; stats_.passes = g.passes; stats_.warpsPerPartition = g.warpsPerPartition;
    return 0;
}

int DeviceDb::search(const unsigned char* query, int Q, int Go, int Ge, const int* matrix, int A, int wantEnd, int mode,
                     const unsigned char* skip, int* scores, int* endQ, int* endT, float* deviceMs) {
    stats_ = SearchStats();
    if (deviceMs) *deviceMs = 0.f;
    if (mode != kModeNW && mode != kModeHW && mode != kModeOV && mode != kModeSW) return OPAL_B200_ERR_MODE;
    if (A <= 0 || A > 254) { set_error("alphabetLength must be in [1, 254]"); return OPAL_B200_ERR_CUDA; }
    // Argument range of the widest pass (reference src/opal.cpp:183-198, 615-630).
    if (Go <= INT_MIN / 2 || INT_MAX / 2 <= Go || Ge <= INT_MIN / 2 || INT_MAX / 2 <= Ge) return OPAL_B200_ERR_OVERFLOW;
    int maxP = INT_MIN, minP = INT_MAX;
    for (int i = 0; i < A * A; i++) {
        if (matrix[i] <= INT_MIN / 2 || INT_MAX / 2 <= matrix[i]) return OPAL_B200_ERR_OVERFLOW;
        maxP = std::max(maxP, matrix[i]);
        minP = std::min(minP, matrix[i]);
    }
    const long long absP = std::max<long long>(std::llabs((long long)maxP), std::llabs((long long)minP));
    const long long gapMax = std::max<long long>(std::llabs((long long)Go), std::llabs((long long)Ge));
    if (Go < 0 || Ge < 0) { set_error("gap penalties must be non-negative"); return OPAL_B200_ERR_OVERFLOW; }
    const bool args16 = absP <= 2048 && gapMax <= 2048;
    const bool args32 = absP < (1 << 28) && gapMax < (1 << 28);
    if (!args32) return OPAL_B200_ERR_OVERFLOW;  // beyond the widths this engine carries (documented deviation)

    if (maxCode_ >= A) { set_error("database holds residue codes >= alphabetLength"); return OPAL_B200_ERR_CUDA; }
    for (int r = 0; r < Q; r++)
        if (query[r] >= A) { set_error("query holds residue codes >= alphabetLength"); return OPAL_B200_ERR_CUDA; }

    // ---- per-target routing
    const bool isSW = mode == kModeSW;
    // NW/HW/OV: every H, E, F of a (Q, T) problem lies in [-(3Go + (Q+T)Ge + |minP|), min(Q,T) maxP + Go].
    auto fits = [&](int T, long long lim) -> bool {
        const long long lo = 3LL * Go + ((long long)Q + T) * Ge + absP;
        const long long hi = (maxP > 0 ? (long long)std::min(Q, T) * maxP : 0) + Go + absP;
        return lo <= lim && hi <= lim;
    };
    std::vector<int> list16, list32;
    bool touched = false;
    for (int p = 0; p < n_; p++) {
        const int i = order_[p];
        if (skip && skip[i]) continue;
        const int T = sortedLen_[p];
        if (T == 0 || Q <= 0) {  // nothing to align: defined as in oracle/opal_oracle.c
            int sc = 0, eq = Q - 1, et = T - 1;
            if (Q > 0 && (mode == kModeNW || mode == kModeHW)) sc = -Go - (Q - 1) * Ge;
            if (isSW) { eq = -1; et = -1; }
            scores[i] = sc;
            if (endQ) endQ[i] = wantEnd ? eq : -1;
            if (endT) endT[i] = wantEnd ? et : -1;
            continue;
        }
        touched = true;
        if (isSW) {
            if (args16) list16.push_back(p);
            else if ((maxP > 0 ? (long long)std::min(Q, T) * maxP : 0) < (1LL << 30)) list32.push_back(p);
            else return OPAL_B200_ERR_OVERFLOW;
        } else {
            if (args16 && fits(T, 28000)) list16.push_back(p);
            else if (fits(T, 1LL << 30)) list32.push_back(p);
            else return OPAL_B200_ERR_OVERFLOW;
        }
    }
    if (!touched) return 0;

    unsigned char* dQuery = nullptr;
    int* dMatrix = nullptr;
    int rc = 0;
    auto body = [&]() -> bool {
        CUDA_TRY(cudaSetDevice(device_));
        if ((size_t)Q + 16 > queryCapacity_) {
            device_release(device_, dQuery_); dQuery_ = nullptr;
            queryCapacity_ = std::max<size_t>(4096, 2 * ((size_t)Q + 16));
            if (!device_alloc(device_, (void**)&dQuery_, queryCapacity_)) return false;
        }
        if (!dMatrix_ && !device_alloc(device_, (void**)&dMatrix_, sizeof(int) * 256 * 256)) return false;
        dQuery = dQuery_; dMatrix = dMatrix_;
        CUDA_TRY(cudaMemcpyAsync(dQuery, query, (size_t)Q, cudaMemcpyHostToDevice, stream_));
        CUDA_TRY(cudaMemcpyAsync(dMatrix, matrix, sizeof(int) * A * A, cudaMemcpyHostToDevice, stream_));
        CUDA_TRY(cudaMemsetAsync(dCounters_, 0, sizeof(int) * 256, stream_));
        int slot = 0;
        CUDA_TRY(cudaEventRecord(evStart_, stream_));
        auto fetch = [&]() -> bool {
            const size_t nInts = sizeof(int) * (size_t)n_;
            CUDA_TRY(cudaMemcpyAsync(hScore_, dScore_, nInts, cudaMemcpyDeviceToHost, stream_));
            if (wantEnd) {
                CUDA_TRY(cudaMemcpyAsync(hEndQ_, dEndQ_, nInts, cudaMemcpyDeviceToHost, stream_));
                CUDA_TRY(cudaMemcpyAsync(hEndT_, dEndT_, nInts, cudaMemcpyDeviceToHost, stream_));
            }
            CUDA_TRY(cudaStreamSynchronize(stream_));
            return true;
        };
        auto publish = [&](const std::vector<int>& list, std::vector<int>* overflowed) {
            for (int p : list) {
                const int i = order_[p];
                const int sc = hScore_[p];
                if (sc == kScoreOverflow || sc == kScoreNone) { if (overflowed) overflowed->push_back(p); else rc = OPAL_B200_ERR_OVERFLOW; continue; }
                scores[i] = sc;
                if (endQ) endQ[i] = (wantEnd && hEndQ_[p] != 0x7fffffff) ? hEndQ_[p] : -1;
                if (endT) endT[i] = (wantEnd && hEndT_[p] != 0x7fffffff) ? hEndT_[p] : -1;
            }
        };
        if (!list16.empty()) {
            rc = run_class(0, list16, dQuery, dMatrix, Q, Go, Ge, A, wantEnd, mode, maxP, &slot);
            if (rc) return rc != OPAL_B200_ERR_CUDA;
            CUDA_TRY(cudaEventRecord(evStop_, stream_));
            if (!fetch()) return false;
            std::vector<int> again;
            publish(list16, &again);
            if (!again.empty()) {
                stats_.rerun32 = (int)again.size();
                std::vector<int> merged(list32.size() + again.size());
                std::merge(list32.begin(), list32.end(), again.begin(), again.end(), merged.begin());
                list32.swap(merged);
            }
        }
        if (!list32.empty()) {
            rc = run_class(1, list32, dQuery, dMatrix, Q, Go, Ge, A, wantEnd, mode, maxP, &slot);
            if (rc) return rc != OPAL_B200_ERR_CUDA;
            CUDA_TRY(cudaEventRecord(evStop_, stream_));
            if (!fetch()) return false;
            publish(list32, nullptr);
        }
        if (deviceMs) CUDA_TRY(cudaEventElapsedTime(deviceMs, evStart_, evStop_));
        return true;
    };
    const bool okb = body();
    if (!okb) return OPAL_B200_ERR_CUDA;
    return rc;
}

// ------------------------------------------------------------------ DPX roofline probe
double measure_dpx_peak(int device, double* threadInstrPerSec, float* msOut) {
    constexpr int ILP = 8;
    const int iters = 4096;
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return 0.0; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 2;
    uint32_t* out = nullptr;
    if (cudaMalloc(&out, sizeof(uint32_t) * blocks * 512) != cudaSuccess) { set_error("cudaMalloc failed"); return 0.0; }
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        dpx_peak_kernel<ILP><<<blocks, 512>>>(out, iters, 12345u + rep);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0) best = std::min(best, ms);
    }
    const cudaError_t e = cudaGetLastError();
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return 0.0; }
    const double instr = 6.0 * ILP * (double)iters * blocks * 512;  // thread-level packed instructions
    const double ips = instr / (best * 1e-3);
    if (threadInstrPerSec) *threadInstrPerSec = ips;
    if (msOut) *msOut = best;
    return ips * 2.0 / 6.0 / 1e9;  // 2 cells per packed instruction, 6 instructions per SW cell pair
}

}  // namespace opalb200
