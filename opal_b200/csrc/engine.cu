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

// Task lengths of one class (longest first) with strided prefix sums, so that the planner can price any
// contiguous range of tasks under any group size in O(1).
struct TaskLens {
    std::vector<int> len;
    std::vector<double> prefix[6];  // prefix[s][k] = sum of len[i] over i = 0, g, 2g, ... < k*g with g = 1 << s
    void build() {
        for (int sh = 0; sh < 6; sh++) {
            const size_t g = (size_t)1 << sh, cnt = (len.size() + g - 1) / g;
            prefix[sh].assign(cnt + 1, 0.0);
            for (size_t k = 0; k < cnt; k++) prefix[sh][k + 1] = prefix[sh][k] + len[k * g];
        }
    }
    // sum of len[i] for i = lo, lo+g, ... < hi  (lo must be a multiple of g = 1 << sh), and how many terms
    double strided(int sh, size_t lo, size_t hi, double* terms) const {
        const size_t g = (size_t)1 << sh, a = lo / g, b = (hi + g - 1) / g;
        *terms = (double)(b - a);
        return prefix[sh][b] - prefix[sh][a];
    }
};

// Picks (G, R, passes, warps per scheduler partition) for a query of Q rows over `taskLens` (one entry
// per task, longest first) on `numSMs` SMs, and returns the estimated cycles in *estCycles.
//
// Timing model (cycles), fitted to B200 runs (profiles/): the integer pipe retires one packed DPX
// instruction per 2 cycles per partition and a cell pair costs ~5.5 of them (11 cycles per row ideal).
// Measured: one warp alone on its partition takes 17.75 R + 104 cycles per step (exposed latencies);
// two warps sharing a partition take 2 (12.8 R + 57) per step each.  Two bounds:
//   throughput:  sum over warp-tasks of steps * stepTime / (partitions * k)
//   tail:        the longest target's steps * stepTime  (it cannot be split across warps)
// Small databases with a long tail (BASELINE configs[1]) are tail-bound and want G = 32 and k = 1;
// large ones are throughput-bound and want fewer threads per target and k = 2.
static bool pick_geometry(int Q, int A, int lanes, const TaskLens& tl, size_t lo, size_t hi, int smemLimit, int numSMs, int mode,
                          bool latencyClass, Geometry* out, double* estCycles) {
    const int planes = lanes == 2 ? 2 : 1;
    const auto& tables = kernel_tables();
    double bestCost = 1e300;
    bool found = false;
    const double tasks = (double)(hi - lo);
    const double maxLen = hi > lo ? tl.len[lo] : 0;
    auto consider = [&](size_t ti, int G, int k, bool forced) {
        const int R = tables[ti].R;
        const int Rpad = rpad_of(R);
        const int rowStride = (G * Rpad + 31) / 32 * 32;
        const size_t smem = (size_t)planes * (A + 1) * rowStride * 4;
        if (smem > (size_t)smemLimit) return;
        const int rows = G * R;
        const int passes = (Q + rows - 1) / rows;
        const int groupsPerWarp = 32 / G;
        // a warp-task lasts as long as its longest group: every groupsPerWarp-th task of the sorted list
        int sh = 0;
        while ((1 << sh) < groupsPerWarp) sh++;
        double warpTasks = 1;
        double warpSteps = tl.strided(sh, lo, hi, &warpTasks);
        warpSteps += warpTasks * (G - 1);
        warpTasks = std::max(1.0, warpTasks);
        // with fewer warp-tasks than resident warps the partitions are not shared k ways
        const double kEff = std::max(1.0, std::min((double)k, std::ceil(warpTasks / (numSMs * 4.0))));
        const double stepTime = kEff < 1.5 ? 17.75 * R + 104.0 : kEff * (12.8 * R + 57.0);
        const double warpsBusy = std::min((double)numSMs * 4 * k, warpTasks);
        const double throughput = warpSteps * stepTime / warpsBusy;
        const double tail = (maxLen + G - 1) * stepTime;
        // every further pass is a kernel of its own (drain, launch, boundary rows through HBM): measured ~4 % each
        const double cost = passes * (std::max(throughput, tail) + 0.15 * std::min(throughput, tail) + 30000.0) *
                            (1.0 + 0.04 * (passes - 1));
        if (forced || cost < bestCost) {
            bestCost = cost;
            found = true;
            out->G = G; out->R = R; out->tableIndex = (int)ti; out->passes = passes; out->Rpad = Rpad;
            out->rowStride = rowStride; out->smemBytes = smem; out->warpsPerPartition = k;
            out->padTop = (mode == kModeNW) ? 0 : passes * rows - Q;
        }
    };
    for (size_t ti = 0; ti < tables.size(); ti++)
        for (int G = latencyClass ? 32 : 1; G <= 32; G *= 2)
            for (int k = 1; k <= (latencyClass ? 1 : kBlockThreads / 128); k *= 2) consider(ti, G, k, false);
    // Development override: OPAL_B200_GEOMETRY="G,R,k" forces a geometry (ignored when it does not fit).
    if (const char* env = getenv("OPAL_B200_GEOMETRY")) {
        int G = 0, R = 0, k = 0;
        if (!latencyClass && sscanf(env, "%d,%d,%d", &G, &R, &k) == 3 && G >= 1 && G <= 32 && (G & (G - 1)) == 0 && k >= 1 &&
            k <= kBlockThreads / 128)
            for (size_t ti = 0; ti < tables.size(); ti++)
                if (tables[ti].R == R) consider(ti, G, k, true);
    }
    if (!found) set_error("alphabet too large for the shared-memory query profile");
    if (estCycles) *estCycles = bestCost;
    return found;
}

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
        if (!stream_acquire(device, &d->stream_) || !event_acquire(device, &d->evStart_) || !event_acquire(device, &d->evStop_) ||
            !event_acquire(device, &d->evFork_)) return false;

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
        // host threads only pay off for large databases (spawning them costs more than copying a few MB)
        int nThreads = total < (32 << 20) ? 1 : (int)std::min<long long>(std::max(1u, std::thread::hardware_concurrency()), total / (8 << 20));
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
        if (!device_alloc(device, (void**)&d->dTaskList_, 2 * nInts)) return false;
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
    event_release(device_, evStart_); event_release(device_, evStop_); event_release(device_, evFork_);
    for (cudaStream_t st : auxStreams_) { cudaStreamSynchronize(st); stream_release(device_, st); }
    for (cudaEvent_t e : auxEvents_) event_release(device_, e);
    stream_release(device_, stream_);
}

bool DeviceDb::ensure_boundary() {
    if (dBndH_) return true;
    const size_t bytes = 2 * sizeof(uint32_t) * (size_t)(totalResidues_ + 64);  // two halves: ping-pong between passes
    if (!device_alloc(device_, (void**)&dBndH_, bytes) || !device_alloc(device_, (void**)&dBndF_, bytes)) return false;
    return true;
}

// One kernel-launch group: tasks that share a precision class and a geometry.
struct DeviceDb::Group {
    int type = 0;            // 0 = Packed16 (tasks are pair indices), 1 = Scalar32 (tasks are sorted-target indices)
    std::vector<int> tasks;  // longest first
    Geometry g;
    size_t smemBytes = 0;    // dynamic shared memory requested (>= g.smemBytes, see exclusive placement below)
    int maxBlocks = 0;
    double estCycles = 0;    // planner's estimate, used to share SMs between concurrent groups
};

// Splits one precision class into launch groups.  A database whose longest targets would keep single warps
// busy long after everything else has finished (the tail bound of pick_geometry) gets a "latency class":
// its longest tasks run with 32 threads per task and one warp per scheduler partition on SMs of their own,
// concurrently with the throughput-oriented bulk group on the remaining SMs.
bool DeviceDb::plan_class(int type, const std::vector<int>& list, int Q, int A, int mode, std::vector<Group>* groups) {
    if (list.empty()) return true;
    const int lanes = type == 0 ? 2 : 1;
    // Packed16 works on the pairs fixed at packing time (targets 2p, 2p+1): a pair runs if either member
    // is wanted; results of unwanted members are simply not published.
    std::vector<int> tasks;
    if (lanes == 2) {
        for (int t : list) if (tasks.empty() || tasks.back() != (t >> 1)) tasks.push_back(t >> 1);
    } else {
        tasks = list;
    }
    TaskLens tl;
    tl.len.resize(tasks.size());
    for (size_t k = 0; k < tasks.size(); k++) tl.len[k] = sortedLen_[lanes == 2 ? 2 * tasks[k] : tasks[k]];
    tl.build();
    const size_t nT = tasks.size();

    Geometry gAll;
    double tAll = 0;
    if (!pick_geometry(Q, A, lanes, tl, 0, nT, smemLimit_, numSMs_, mode, false, &gAll, &tAll)) return false;
    // candidate splits: the m longest tasks form the latency class on smL SMs
    double bestT = tAll;
    size_t bestM = 0;
    int bestSm = 0;
    Geometry bestL, bestB;
    if (!getenv("OPAL_B200_NO_SPLIT") && !getenv("OPAL_B200_GEOMETRY")) {
        for (size_t m = 32; m * 2 <= nT && m <= 4096; m *= 2) {  // multiples of 32 keep every group size aligned
            for (int div = 1; div <= 4; div *= 2) {
                const int smL = (int)std::min<size_t>((m + 4 * div - 1) / (4 * div), (size_t)numSMs_ / 2);
                if (smL < 1) continue;
                Geometry gL, gB;
                double tL = 0, tB = 0;
                if (!pick_geometry(Q, A, lanes, tl, 0, m, smemLimit_, smL, mode, true, &gL, &tL)) continue;
                if (!pick_geometry(Q, A, lanes, tl, m, nT, smemLimit_, numSMs_ - smL, mode, false, &gB, &tB)) continue;
                const double t = std::max(tL, tB) + 3000.0;  // a second launch is not free
                if (t < bestT * 0.97) { bestT = t; bestM = m; bestSm = smL; bestL = gL; bestB = gB; }
            }
        }
    }
    auto add = [&](size_t lo, size_t hi, const Geometry& g, int maxBlocks, size_t otherSmem, double est) {
        Group grp;
        grp.type = type;
        grp.estCycles = est;
        grp.tasks.assign(tasks.begin() + lo, tasks.begin() + hi);
        grp.g = g;
        grp.maxBlocks = maxBlocks;
        grp.smemBytes = g.smemBytes;
        // exclusive placement: ask for enough shared memory that a block of the other group cannot join this SM
        if (otherSmem > 0 && g.smemBytes + otherSmem <= (size_t)smemLimit_)
            grp.smemBytes = std::min<size_t>((size_t)smemLimit_, (size_t)smemLimit_ - otherSmem + 1024);
        groups->push_back(std::move(grp));
    };
    if (bestM == 0) add(0, tasks.size(), gAll, numSMs_, 0, tAll);
    else {
        add(0, bestM, bestL, bestSm, bestB.smemBytes, bestT);   // launched first: takes its SMs
        // every block's first tasks are assigned statically (longest first), so the bulk grid must be fully
        // resident from the start: one block per SM the latency class leaves free
        add(bestM, tasks.size(), bestB, numSMs_ - bestSm, 0, bestT);
    }
    return true;
}

// Launches every pass of one group on `stream`.
bool DeviceDb::launch_group(const Group& grp, int* taskListDevice, cudaStream_t stream, const unsigned char* dQuery,
                            const int* dMatrix, int Q, int Go, int Ge, int A, int wantEnd, int mode, int maxScore, int* launchSlot) {
    const Geometry& g = grp.g;
    const int type = grp.type;
    if (g.passes > 1 && !ensure_boundary()) return false;
    if (*launchSlot + g.passes > 256) { set_error("too many passes"); return false; }
    // contiguous task ranges need no list in device memory
    bool contiguous = true;
    for (size_t k = 1; k < grp.tasks.size() && contiguous; k++) contiguous = grp.tasks[k] == grp.tasks[k - 1] + 1;
    if (!contiguous)
        CUDA_TRY(cudaMemcpyAsync(taskListDevice, grp.tasks.data(), sizeof(int) * grp.tasks.size(), cudaMemcpyHostToDevice, stream));
    // SW score+end at 16 bits uses the key-tracking flavor: exact below fastEndLimit, and warps that meet a
    // larger score sweep their tasks again with the exact per-cell tracking inside the same kernel.
    const int fastEndLimit = (32768 >> kRowBits) - std::max(maxScore, 0) - 1;
    const bool fastEnd = mode == kModeSW && wantEnd && type == 0 && g.R <= (1 << kRowBits) && fastEndLimit >= 64 &&
                         !getenv("OPAL_B200_EXACT_END");
    const int flavor = (mode == kModeSW) ? (wantEnd ? (fastEnd ? kFlavorSWEndFast : kFlavorSWEnd) : kFlavorSWScore) : kFlavorGlobal;
    const void* fn = kernel_tables()[g.tableIndex].fn[type * 4 + flavor];
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grp.smemBytes));
    for (int pass = 0; pass < g.passes; pass++) {
        SearchParams p;
        memset(&p, 0, sizeof(p));
        p.query = dQuery; p.matrix = dMatrix; p.Q = Q; p.A = A; p.gapOpen = Go; p.gapExt = Ge; p.mode = mode;
        p.wantEnd = wantEnd;
        p.G = g.G; p.rowBase = pass * g.G * g.R; p.padTop = g.padTop; p.pass = pass; p.numPasses = g.passes;
        p.rowStride = g.rowStride; p.Rpad = g.Rpad;
        p.residues = dResidues_; p.offsets = dOffsets_; p.lengths = dLengths_;
        p.pairStream = dPairStream_; p.pairOffsets = dPairOffsets_; p.numTargets = n_;
        p.taskList = contiguous ? nullptr : taskListDevice;
        p.taskBase = contiguous && !grp.tasks.empty() ? grp.tasks[0] : 0;
        p.numTasks = (int)grp.tasks.size();
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
        const long long warpsNeeded = ((long long)grp.tasks.size() * g.G + 31) / 32;
        const int blocks = (int)std::max<long long>(1, std::min<long long>(grp.maxBlocks, (warpsNeeded + warpsPerBlock - 1) / warpsPerBlock));
        CUDA_TRY(cudaLaunchKernel(fn, dim3(blocks), dim3(32 * warpsPerBlock), args, grp.smemBytes, stream));
        stats_.kernelLaunches++;
    }
    stats_.G = g.G; stats_.R = g.R; stats_.passes = g.passes; stats_.warpsPerPartition = g.warpsPerPartition;
    return true;
}

// Runs the given classes concurrently: every launch group gets a stream of its own (the first one the
// database's main stream), all of them ordered after `ready` and joined back into the main stream.
int DeviceDb::run_classes(const std::vector<std::pair<int, const std::vector<int>*>>& classes, const unsigned char* dQuery,
                          const int* dMatrix, int Q, int Go, int Ge, int A, int wantEnd, int mode, int maxScore, int* launchSlot) {
    std::vector<Group> groups;
    for (auto& c : classes)
        if (!plan_class(c.first, *c.second, Q, A, mode, &groups)) return OPAL_B200_ERR_CUDA;
    if (groups.empty()) return 0;
    stats_.groups = (int)groups.size();
    // Concurrent groups own disjoint SMs (grids are persistent, one block per SM, and every block must be
    // resident from the start because first tasks are assigned statically).  Small groups -- the latency
    // class, a handful of 32-bit targets -- get the blocks they can fill and are launched first; the
    // remaining SMs are shared by the large groups in proportion to their estimated work.
    if (groups.size() > 1) {
        auto wanted = [&](const Group& g) {
            const int wpb = 4 * g.g.warpsPerPartition;
            const long long warps = ((long long)g.tasks.size() * g.g.G + 31) / 32;
            return (int)std::max<long long>(1, std::min<long long>(g.maxBlocks, (warps + wpb - 1) / wpb));
        };
        std::stable_sort(groups.begin(), groups.end(), [&](const Group& a, const Group& b) { return wanted(a) < wanted(b); });
        std::vector<int> want(groups.size());
        std::vector<char> big(groups.size(), 0);
        for (size_t i = 0; i < groups.size(); i++) want[i] = wanted(groups[i]);
        int reserved = 0;
        double bigWork = 0;
        for (size_t i = 0; i < groups.size(); i++) {
            if (want[i] * 4 <= numSMs_ && reserved + want[i] <= numSMs_ / 2) { groups[i].maxBlocks = want[i]; reserved += want[i]; }
            else { big[i] = 1; bigWork += std::max(1.0, groups[i].estCycles * want[i]); }
        }
        const int left = numSMs_ - reserved;
        for (size_t i = 0; i < groups.size(); i++) {
            if (!big[i]) continue;
            const int share = std::max(1, (int)(left * std::max(1.0, groups[i].estCycles * want[i]) / bigWork));
            groups[i].maxBlocks = std::min(share, want[i]);
        }
    }
    if (getenv("OPAL_B200_TRACE"))
        for (const Group& g : groups)
            fprintf(stderr, "[opal-b200] group type=%d tasks=%zu longest=%d G=%d R=%d k=%d passes=%d blocks<=%d est=%.0f kcycles\n", g.type,
                    g.tasks.size(), g.tasks.empty() ? 0 : sortedLen_[g.type == 0 ? 2 * g.tasks[0] : g.tasks[0]], g.g.G, g.g.R,
                    g.g.warpsPerPartition, g.g.passes, g.maxBlocks, g.estCycles / 1e3);
    auto body = [&]() -> bool {
        size_t listOffset = 0;
        if (!startRecorded_) { CUDA_TRY(cudaEventRecord(evStart_, stream_)); startRecorded_ = true; }  // planning is host work: keep it outside the device window
        if (groups.size() > 1) CUDA_TRY(cudaEventRecord(evFork_, stream_));
        for (size_t gi = 0; gi < groups.size(); gi++) {
            cudaStream_t st = stream_;
            if (gi > 0) {
                while (auxStreams_.size() < gi) {
                    cudaStream_t s2; cudaEvent_t e2;
                    if (!stream_acquire(device_, &s2) || !event_acquire(device_, &e2)) return false;
                    auxStreams_.push_back(s2); auxEvents_.push_back(e2);
                }
                st = auxStreams_[gi - 1];
                CUDA_TRY(cudaStreamWaitEvent(st, evFork_, 0));
            }
            if (!launch_group(groups[gi], dTaskList_ + listOffset, st, dQuery, dMatrix, Q, Go, Ge, A, wantEnd, mode, maxScore, launchSlot))
                return false;
            listOffset += groups[gi].tasks.size();
            if (gi > 0) {
                CUDA_TRY(cudaEventRecord(auxEvents_[gi - 1], st));
                CUDA_TRY(cudaStreamWaitEvent(stream_, auxEvents_[gi - 1], 0));
            }
        }
        return true;
    };
    return body() ? 0 : OPAL_B200_ERR_CUDA;
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
    // The 16-bit class works on whole pairs (sorted targets 2p, 2p+1) and the two classes of NW/HW/OV run
    // concurrently, so a pair is never split between them: if one member needs 32 bits, both go there.
    if (!isSW && !list32.empty() && !list16.empty()) {
        std::vector<char> wide((size_t)numPairs_, 0);
        for (int p : list32) wide[p >> 1] = 1;
        std::vector<int> keep16, merged32;
        for (int p : list16) (wide[p >> 1] ? merged32 : keep16).push_back(p);
        if (!merged32.empty()) {
            std::vector<int> all(list32.size() + merged32.size());
            std::merge(list32.begin(), list32.end(), merged32.begin(), merged32.end(), all.begin());
            list32.swap(all);
            list16.swap(keep16);
        }
    }

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
        startRecorded_ = false;
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
        // NW/HW/OV classes are independent (routed a priori) and run concurrently; SW's 32-bit class is the
        // re-run of what overflowed 16 bits and has to follow it.
        std::vector<std::pair<int, const std::vector<int>*>> first;
        if (!list16.empty()) first.push_back({0, &list16});
        if (!isSW && !list32.empty()) first.push_back({1, &list32});
        if (!first.empty()) {
            rc = run_classes(first, dQuery, dMatrix, Q, Go, Ge, A, wantEnd, mode, maxP, &slot);
            if (rc) return rc != OPAL_B200_ERR_CUDA;
            CUDA_TRY(cudaEventRecord(evStop_, stream_));
            if (!fetch()) return false;
            std::vector<int> again;
            publish(list16, isSW ? &again : nullptr);
            if (!isSW) publish(list32, nullptr);
            if (!again.empty()) {
                stats_.rerun32 = (int)again.size();
                std::vector<int> merged(list32.size() + again.size());
                std::merge(list32.begin(), list32.end(), again.begin(), again.end(), merged.begin());
                list32.swap(merged);
            }
        }
        if (isSW && !list32.empty()) {
            std::vector<std::pair<int, const std::vector<int>*>> second = {{1, &list32}};
            rc = run_classes(second, dQuery, dMatrix, Q, Go, Ge, A, wantEnd, mode, maxP, &slot);
            if (rc) return rc != OPAL_B200_ERR_CUDA;
            CUDA_TRY(cudaEventRecord(evStop_, stream_));
            if (!fetch()) return false;
            publish(list32, nullptr);
        }
        if (deviceMs && startRecorded_) CUDA_TRY(cudaEventElapsedTime(deviceMs, evStart_, evStop_));
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
