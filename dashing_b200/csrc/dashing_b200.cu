// dashing_b200.cu — the C ABI (include/dashing_b200.h) over the sm_100a kernels in sketch.cuh / dist.cuh.
// Host orchestration only: device memory, streams, H2D/D2H, launch geometry.  No arithmetic of the
// hot paths runs on the CPU here and there is no fallback: without a CUDA device every compute
// entry point returns DB200_ENODEV.
#include "common.cuh"
#include "estimators.cuh"
#include "sketch.cuh"
#include "dist.cuh"
#include "setops.cuh"
#include "fasta.cuh"

#include <algorithm>
#include <cstring>
#include <sched.h>
#include <chrono>
#include <mutex>
#include <thread>
#include <condition_variable>
#include <functional>
#include <vector>
#include <memory>

namespace db200 {

static thread_local std::string t_err;
static thread_local double g_dbg_pack_ms = 0, g_dbg_wait_ms = 0;   // DB200_DEBUG_UPLOAD accounting
std::atomic<uint64_t> g_kernel_launches{0};

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_err = buf;
}

static int g_num_sms(int device) {
    static int cache[64] = {0};
    if (device < 64 && cache[device]) return cache[device];
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, phys_of(device));
    if (device < 64) cache[device] = v;
    return v;
}

// ---------------------------------------------------------------------------------------------
// packed genome store
// ---------------------------------------------------------------------------------------------
} // namespace db200

struct db200_packed_genomes {
    int device = 0, k = 0;
    uint64_t nbases = 0, nblk = 0, ngenomes = 0, kmers = 0;
    uint32_t nitems = 0;
    std::vector<uint32_t> group_item_begin;   // items of genome group g: [group_item_begin[g], group_item_begin[g+1])
    std::vector<uint64_t> group_end;          // one past the last base of group g
    db200::DevBuf bases2, nb, st, items, counter, starts;
    // work counters: [0, 8) belong to the pipelined genome groups of one pack call, [8, 16) rotate over the launches of
    // db200_sketch_packed_dev so that sketches of the same store on different streams never share (or reset) a counter
    mutable std::atomic<uint32_t> launch_seq{0};
    uint64_t host_packed_chunks = 0, ascii_chunks = 0;   // how the last pack call split its chunks between the two upload routes
};

extern "C" void db200_hostpack(const uint8_t *ascii, size_t nbases, uint32_t *codes, uint16_t *valid);   // hostpack.cpp

namespace db200 {
// CPUs this process may really use: the affinity mask capped by the cgroup CPU quota (the GPU boxes expose 128 logical CPUs
// but grant 16 CPUs of time; oversubscribing a quota is far slower than matching it).
static unsigned usable_cpus() {
    unsigned n = std::max(1u, std::thread::hardware_concurrency());
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = std::max(1, CPU_COUNT(&set));
    if (std::FILE *f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char q[64] = {0};
        unsigned long long per = 0;
        if (std::fscanf(f, "%63s %llu", q, &per) == 2 && std::strcmp(q, "max") != 0 && per > 0) {
            const unsigned long long quota = std::strtoull(q, nullptr, 10);
            if (quota > 0) n = std::min<unsigned>(n, (unsigned)std::max<unsigned long long>(1, (quota + per - 1) / per));
        }
        std::fclose(f);
    }
    return n;
}

// Host -> device copies from PAGEABLE memory.  cudaMemcpyAsync would stage them through the driver's own bounce buffer with one
// thread (~10 GB/s measured on the B200 boxes, against ~55 GB/s for page-locked sources); page-locking the caller's buffer
// costs ~0.5 s per GB.  Instead the library keeps a small ring of page-locked buffers per device and fills it with several
// host threads, so the DMA engine sees page-locked sources at memory-copy speed.  Page-locked sources are passed through.
class CopyPool {
public:
    explicit CopyPool(unsigned n) : nworkers_(n) {
        for (unsigned i = 0; i < n; ++i) th_.emplace_back([this, i] { run(i); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    // dst[0, n) = src[0, n), split over the workers and the calling thread
    void copy(char *dst, const char *src, size_t n) {
        if (n < (size_t(1) << 20) || nworkers_ == 0) { std::memcpy(dst, src, n); return; }
        parallel(n, 4096, [dst, src](size_t b, size_t e) { std::memcpy(dst + b, src + b, e - b); });
    }
    // fn(begin, end) over [0, n) in nworkers + 1 contiguous parts whose boundaries are multiples of `align`
    void parallel(size_t n, size_t align, const std::function<void(size_t, size_t)> &fn) {
        if (nworkers_ == 0 || n <= align) { fn(0, n); return; }
        const size_t parts = nworkers_ + 1, each = ((n + parts - 1) / parts + align - 1) / align * align;
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; each_ = each; pending_ = nworkers_; ++gen_;
        }
        cv_.notify_all();
        const size_t off = each * nworkers_;
        if (off < n) fn(off, n);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
    unsigned workers() const { return nworkers_; }
private:
    void run(unsigned i) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            if (stop_) return;
            const std::function<void(size_t, size_t)> *fn = fn_; const size_t n = n_, each = each_;
            lk.unlock();
            const size_t off = each * i;
            if (off < n) (*fn)(off, off + std::min(each, n - off));
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    unsigned nworkers_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t, size_t)> *fn_ = nullptr;
    size_t n_ = 0, each_ = 0;
    unsigned pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct HostStager {
    static constexpr int NSLOT = 4;
    static constexpr size_t SLOT = size_t(16) << 20;
    char *slot[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t freed[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    bool used[NSLOT] = {false, false, false, false};
    unsigned next = 0;
    std::unique_ptr<CopyPool> pool;
    int init() {
        if (slot[0]) return DB200_OK;
        for (int i = 0; i < NSLOT; ++i) {
            DB200_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&slot[i]), SLOT, cudaHostAllocDefault));
            DB200_CUDA(cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming));
        }
        const char *e = std::getenv("DB200_COPY_THREADS");
        int n = e ? std::atoi(e) : 8;
        n = std::max(1, std::min(n, 64));
        pool.reset(new CopyPool((unsigned)n - 1));
        return DB200_OK;
    }
    static bool page_locked(const void *p) {
        cudaPointerAttributes a;
        const bool ok = cudaPointerGetAttributes(&a, p) == cudaSuccess && (a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged || a.type == cudaMemoryTypeDevice);
        cudaGetLastError();
        return ok;
    }
    // Device -> host into possibly pageable memory, ordered after the work already queued on `cs`.  Page-locked destinations get
    // one asynchronous copy; pageable ones are filled from the bounce buffers by the copy threads while the next piece is in
    // flight (the call then returns with the data in place).
    int download(char *dst, const void *src, size_t n, cudaStream_t cs) {
        if (n == 0) return DB200_OK;
        if (page_locked(dst)) { DB200_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, cs)); return DB200_OK; }
        DB200_TRY(init());
        size_t poff = 0, plen = 0;
        unsigned ps = 0;
        bool have = false;
        for (size_t off = 0; off < n; off += SLOT) {
            const size_t len = std::min(SLOT, n - off);
            const unsigned sl = next++ % NSLOT;
            if (used[sl]) DB200_CUDA(cudaEventSynchronize(freed[sl]));
            DB200_CUDA(cudaMemcpyAsync(slot[sl], static_cast<const char *>(src) + off, len, cudaMemcpyDeviceToHost, cs));
            DB200_CUDA(cudaEventRecord(freed[sl], cs));
            used[sl] = true;
            if (have) { DB200_CUDA(cudaEventSynchronize(freed[ps])); pool->copy(dst + poff, slot[ps], plen); }
            poff = off; plen = len; ps = sl; have = true;
        }
        if (have) { DB200_CUDA(cudaEventSynchronize(freed[ps])); pool->copy(dst + poff, slot[ps], plen); }
        return DB200_OK;
    }
    // Enqueues dst[0, n) = src[0, n) on stream `cs`.  Returns once every byte of `src` has been read or handed to the DMA
    // engine from a page-locked source (the caller may not reuse page-locked sources before the stream has drained).
    int upload(void *dst, const char *src, size_t n, cudaStream_t cs) {
        if (n == 0) return DB200_OK;
        if (page_locked(src)) { DB200_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, cs)); return DB200_OK; }
        DB200_TRY(init());
        for (size_t off = 0; off < n; off += SLOT) {
            const size_t len = std::min(SLOT, n - off);
            const unsigned s = next++ % NSLOT;
            if (used[s]) DB200_CUDA(cudaEventSynchronize(freed[s]));
            pool->copy(slot[s], src + off, len);
            DB200_CUDA(cudaMemcpyAsync(static_cast<char *>(dst) + off, slot[s], len, cudaMemcpyHostToDevice, cs));
            DB200_CUDA(cudaEventRecord(freed[s], cs));
            used[s] = true;
        }
        return DB200_OK;
    }
};

// ASCII upload pipeline state (two device staging buffers, a copy stream and events), kept across calls:
// cudaMalloc/cudaFree of multi-GB buffers costs more than the transfers they serve.
struct Uploader {
    static constexpr int NSTAGE = 4;           // device staging buffers for ASCII chunks = ASCII copies that may be queued at once
    DevBuf stage[NSTAGE];
    cudaStream_t cs = nullptr;
    cudaEvent_t copied[NSTAGE] = {}, packed[NSTAGE] = {};
    static constexpr int NGROUP = 8;           // only the LAST group's sketch is exposed after the upload: keep it short
    HostStager stager;                         // pageable sources go through page-locked bounce buffers filled by several threads
    cudaStream_t ss = nullptr;                 // sketch stream: group g is sketched while later groups are still uploading
    cudaEvent_t group_packed[NGROUP] = {};
    // host-packed share of a batch (hostpack.cpp): page-locked slots of 2-bit codes + validity for one 64 Mbase chunk each,
    // filled by the pack threads and copied straight into the store on a second copy stream
    static constexpr int NPSLOT = 3;
    static constexpr uint64_t CHUNK = 64ull << 20;     // bases per chunk (a multiple of 64)
    char *pslot[NPSLOT] = {nullptr, nullptr, nullptr};
    cudaEvent_t pslot_done[NPSLOT] = {nullptr, nullptr, nullptr};
    bool pslot_used[NPSLOT] = {false, false, false};
    std::atomic<bool> pslot_inflight[NPSLOT];   // a packed copy has been enqueued from this slot (read by the feeder thread)
    cudaStream_t cs2 = nullptr;
    std::unique_ptr<CopyPool> packers;
    int init_hostpack() {
        if (pslot[0]) return DB200_OK;
        for (auto &f : pslot_inflight) f.store(false);
        DB200_CUDA(cudaStreamCreateWithFlags(&cs2, cudaStreamNonBlocking));
        for (int i = 0; i < NPSLOT; ++i) {
            DB200_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&pslot[i]), CHUNK / 16 * 6, cudaHostAllocDefault));
            DB200_CUDA(cudaEventCreateWithFlags(&pslot_done[i], cudaEventDisableTiming));
        }
        const char *e = std::getenv("DB200_PACK_THREADS");
        int n = e ? std::atoi(e) : (int)usable_cpus();
        n = std::max(1, std::min(n, 64));
        packers.reset(new CopyPool((unsigned)n - 1));
        return DB200_OK;
    }
    int init() {
        if (cs) return DB200_OK;
        DB200_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        DB200_CUDA(cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking));
        for (int i = 0; i < NSTAGE; ++i) {
            DB200_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
            DB200_CUDA(cudaEventCreateWithFlags(&packed[i], cudaEventDisableTiming));
        }
        for (auto &e : group_packed) DB200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return DB200_OK;
    }
};
// When given to pack_genomes_impl, every genome group is sketched as soon as its last chunk has been packed.
struct PipelinedSketch { int p, canon; uint8_t *d_regs; };
static int sketch_launch(const db200_packed_genomes *pg, int p, int canon, uint8_t *d_regs, cudaStream_t stream, uint32_t item_begin,
                         uint32_t item_count, int counter_idx);
} // namespace db200

namespace db200 {

static uint64_t count_kmers(const uint64_t *rec_offsets, uint64_t nrecords, int k) {
    // work items of the metric: sum over records of max(0, len - k + 1)   
    uint64_t t = 0;
    for (uint64_t r = 0; r < nrecords; ++r) {
        const uint64_t len = rec_offsets[r + 1] - rec_offsets[r];
        if (len >= (uint64_t)k) t += len - k + 1;
    }
    return t;
}

static int pack_genomes_impl(int device, const char *bases, const uint64_t *rec_offsets, uint64_t nrecords,
                             const uint64_t *genome_rec_begin, uint64_t ngenomes, int k, db200_packed_genomes *pg,
                             Uploader &up, cudaStream_t stream, const PipelinedSketch *ps = nullptr) {
    const uint64_t base0 = nrecords ? rec_offsets[0] : 0;
    const uint64_t T = nrecords ? rec_offsets[nrecords] - base0 : 0;
    pg->device = device; pg->k = k; pg->nbases = T; pg->ngenomes = ngenomes;
    pg->nblk = (T + 63) / 64 + 1;  // +1: an all-invalid guard block (idle lanes and the predecessor of block 0 read it)
    pg->kmers = count_kmers(rec_offsets, nrecords, k);
    DB200_TRY(pg->bases2.reserve(pg->nblk * 16));
    DB200_TRY(pg->nb.reserve(pg->nblk * 8));
    DB200_TRY(pg->st.reserve(pg->nblk * 8));
    DB200_TRY(pg->counter.reserve(64));
    DB200_CUDA(cudaMemsetAsync(pg->st.ptr, 0, pg->nblk * 8, stream));
    // the pack kernel writes whole 16-base groups; only the tail of the last block and the guard block need zeroing
    const uint64_t full_blk = T / 64;
    DB200_CUDA(cudaMemsetAsync(pg->nb.as<uint64_t>() + full_blk, 0, (pg->nblk - full_blk) * 8, stream));
    DB200_CUDA(cudaMemsetAsync(pg->bases2.as<uint4>() + full_blk, 0, (pg->nblk - full_blk) * 16, stream));

    // record starts
    std::vector<uint64_t> starts;
    starts.reserve(nrecords);
    for (uint64_t r = 0; r < nrecords; ++r) {
        const uint64_t pos = rec_offsets[r] - base0;
        if (pos < T && rec_offsets[r + 1] > rec_offsets[r]) starts.push_back(pos);
    }
    if (!starts.empty()) {
        DB200_TRY(pg->starts.reserve(starts.size() * 8));
        DB200_CUDA(cudaMemcpyAsync(pg->starts.ptr, starts.data(), starts.size() * 8, cudaMemcpyHostToDevice, stream));
        mark_starts_kernel<<<(unsigned)((starts.size() + 255) / 256), 256, 0, stream>>>(pg->starts.as<uint64_t>(), starts.size(), pg->st.as<uint32_t>());
        DB200_LAUNCHED();
    }
    // work items: every genome is cut into equal chunks of ~1 Mbase (at least ~16 items per SM overall).  Genomes form up to
    // NGROUP contiguous groups of similar size (the unit of upload/compute overlap); inside a group items are ordered
    // chunk-major so that by the time chunk c+1 of a genome starts, chunk c has been merged into HBM and seeds the staged
    // registers (fewer updates).
    std::vector<SketchItem> items;
    pg->group_item_begin.assign(1, 0u);
    pg->group_end.clear();
    {
        const uint64_t target_items = (uint64_t)g_num_sms(device) * 16;
        uint64_t chunk = T / std::max<uint64_t>(target_items, 1);
        chunk = std::min<uint64_t>(std::max<uint64_t>(chunk, 1ull << 16), 1ull << 20);
        const int ngroups = ps ? Uploader::NGROUP : 1;
        uint64_t g0 = 0;
        for (int grp = 0; grp < ngroups && g0 < ngenomes; ++grp) {
            uint64_t g1 = ngenomes;
            if (grp + 1 < ngroups) {
                const uint64_t want = T * (uint64_t)(grp + 1) / (uint64_t)ngroups;
                g1 = g0 + 1;
                while (g1 < ngenomes && rec_offsets[genome_rec_begin[g1]] - base0 < want) ++g1;
            }
            std::vector<std::pair<uint32_t, SketchItem>> tmp;
            for (uint64_t g = g0; g < g1; ++g) {
                const uint64_t gs = rec_offsets[genome_rec_begin[g]] - base0, ge = rec_offsets[genome_rec_begin[g + 1]] - base0;
                if (ge <= gs) continue;
                const uint64_t nch = (ge - gs + chunk - 1) / chunk;
                const uint64_t step = (((ge - gs + nch - 1) / nch) + 63) & ~63ull;
                uint32_t ci = 0;
                for (uint64_t s = gs; s < ge; s += step, ++ci) tmp.push_back({ci, SketchItem{s, std::min(ge, s + step), (uint32_t)g, 0}});
            }
            std::stable_sort(tmp.begin(), tmp.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
            for (auto &t : tmp) items.push_back(t.second);
            pg->group_item_begin.push_back((uint32_t)items.size());
            pg->group_end.push_back(rec_offsets[genome_rec_begin[g1]] - base0);
            g0 = g1;
        }
    }
    pg->nitems = (uint32_t)items.size();
    if (!items.empty()) {
        DB200_TRY(pg->items.reserve(items.size() * sizeof(SketchItem)));
        DB200_CUDA(cudaMemcpyAsync(pg->items.ptr, items.data(), items.size() * sizeof(SketchItem), cudaMemcpyHostToDevice, stream));
    }
    DB200_CUDA(cudaMemsetAsync(pg->counter.ptr, 0, 64, stream));
    if (ps) DB200_CUDA(cudaMemsetAsync(ps->d_regs, 0, ngenomes << ps->p, stream));
    // ASCII -> packed store in 64-base-aligned chunks, by two routes that share the host->device link:
    //  (A) the ASCII bytes go up through two device staging buffers and pack_kernel packs them (1 byte per base on the link;
    //      the pack kernel of chunk c overlaps the copy of the next chunk);
    //  (B) host threads pack the chunk to 2-bit codes + validity (hostpack.cpp) into a page-locked slot that is copied
    //      straight into the store (0.375 bytes per base on the link).
    // With a page-locked source the chunks are consumed from BOTH ends — (B) from the front as fast as the host cores pack,
    // (A) from the back whenever fewer than two ASCII copies are queued — so the link never idles and the split follows
    // the measured speeds.  A pageable source would have to be copied by host threads anyway (HostStager): it is packed
    // instead, all of it.  DB200_HOST_PACK=0 keeps everything on route (A).
    // `bases` may already live in device memory (unified addressing): pack straight from it when aligned
    cudaPointerAttributes pat;
    const bool on_dev = T && cudaPointerGetAttributes(&pat, bases) == cudaSuccess && pat.type == cudaMemoryTypeDevice &&
                        ((reinterpret_cast<uintptr_t>(bases) + base0) & 15) == 0;
    cudaGetLastError();
    const char *hpenv = std::getenv("DB200_HOST_PACK");
    const bool host_pack = !on_dev && T && !(hpenv && hpenv[0] == '0');
    const bool src_locked = !on_dev && T && HostStager::page_locked(bases);
    // unit of the schedule: 64 Mbases on the ASCII-only route; 16 Mbases when both routes share the link, so that a host-packed
    // copy never queues behind more than two short ASCII copies (the first version used 64 MB units and four queued ASCII
    // copies: the packed copies waited 35 ms of a 66 ms batch behind them and only 28 of 75 chunks got packed)
    uint64_t CH = host_pack ? Uploader::CHUNK / 4 : Uploader::CHUNK;
    if (const char *cenv = std::getenv("DB200_UPLOAD_CHUNK")) {     // testing knob: many small chunks exercise the two-ended schedule
        const uint64_t v = std::strtoull(cenv, nullptr, 10) / 4096 * 4096;
        if (v) CH = std::min(v, CH);
    }
    const uint64_t nchunks = (T + CH - 1) / CH;
    if (!on_dev && T) {
        DB200_TRY(up.init());
        if (!host_pack || src_locked)
            for (int i = 0; i < (host_pack ? Uploader::NSTAGE : 2); ++i) DB200_TRY(up.stage[i].reserve(std::min<uint64_t>(CH, T)));
        if (host_pack) DB200_TRY(up.init_hostpack());
        // the copy streams must not run ahead of work already queued on `stream` (memsets above, kernels still reading the staging buffers)
        DB200_CUDA(cudaEventRecord(up.packed[0], stream));
        DB200_CUDA(cudaStreamWaitEvent(up.cs, up.packed[0], 0));
        if (host_pack) DB200_CUDA(cudaStreamWaitEvent(up.cs2, up.packed[0], 0));
    }
    std::vector<char> chunk_done(nchunks, 0), group_launched(pg->group_end.size(), 0);
    // genome groups all of whose chunks have been enqueued start sketching on their own stream
    auto launch_ready_groups = [&]() -> int {
        if (!ps) return DB200_OK;
        uint64_t gb = 0;
        for (size_t g = 0; g < pg->group_end.size(); ++g) {
            const uint64_t ge = pg->group_end[g];
            if (!group_launched[g]) {
                bool ready = true;
                for (uint64_t c = gb / CH; ready && c < nchunks && c * CH < std::max(ge, gb + 1); ++c) ready = chunk_done[c] != 0;
                if (ready) {
                    group_launched[g] = 1;
                    const uint32_t ib = pg->group_item_begin[g], ie = pg->group_item_begin[g + 1];
                    if (ie > ib) {
                        DB200_TRY(up.init());
                        DB200_CUDA(cudaEventRecord(up.group_packed[g], stream));
                        DB200_CUDA(cudaStreamWaitEvent(up.ss, up.group_packed[g], 0));
                        DB200_TRY(sketch_launch(pg, ps->p, ps->canon, ps->d_regs, up.ss, ib, ie - ib, (int)g));
                    }
                }
            }
            gb = ge;
        }
        return DB200_OK;
    };
    std::atomic<bool> stage_used[Uploader::NSTAGE];
    for (auto &f : stage_used) f.store(false);
    uint64_t ascii_seq = 0;
    // route (A): chunk c through staging buffer b
    auto enqueue_ascii = [&](uint64_t c, int b) -> int {
        const uint64_t off = c * CH, len = std::min<uint64_t>(CH, T - off);
        const uint64_t ngroups = (len + 15) / 16;
        const uint8_t *src;
        if (on_dev) {
            src = reinterpret_cast<const uint8_t *>(bases) + base0 + off;
        } else {
            if (stage_used[b]) DB200_CUDA(cudaStreamWaitEvent(up.cs, up.packed[b], 0));  // staging buffer free again
            DB200_TRY(up.stager.upload(up.stage[b].ptr, bases + base0 + off, len, up.cs));
            DB200_CUDA(cudaEventRecord(up.copied[b], up.cs));
            DB200_CUDA(cudaStreamWaitEvent(stream, up.copied[b], 0));
            src = up.stage[b].as<uint8_t>();
        }
        pack_kernel<<<(unsigned)((ngroups + 255) / 256), 256, 0, stream>>>(src, len, pg->bases2.as<uint32_t>() + off / 16,
                                                                          pg->nb.as<uint16_t>() + off / 16, ngroups);
        DB200_LAUNCHED();
        if (!on_dev) { DB200_CUDA(cudaEventRecord(up.packed[b], stream)); stage_used[b] = true; }
        return DB200_OK;
    };
    // route (B): chunk c packed on the host into page-locked slot `sl`
    uint64_t pack_seq = 0;
    auto enqueue_packed = [&](uint64_t c, uint64_t nunits) -> int {
        const uint64_t off = c * CH, len = std::min<uint64_t>(nunits * CH, T - off);
        const uint64_t ngroups = (len + 15) / 16;
        const int sl = (int)(pack_seq++ % Uploader::NPSLOT);
        const auto tw0 = std::chrono::steady_clock::now();
        if (up.pslot_used[sl]) DB200_CUDA(cudaEventSynchronize(up.pslot_done[sl]));
        const auto tw1 = std::chrono::steady_clock::now();
        uint32_t *hc = reinterpret_cast<uint32_t *>(up.pslot[sl]);
        uint16_t *hv = reinterpret_cast<uint16_t *>(up.pslot[sl] + Uploader::CHUNK / 16 * 4);
        const uint8_t *srcb = reinterpret_cast<const uint8_t *>(bases) + base0 + off;
        up.packers->parallel(len, 4096, [=](size_t b0, size_t b1) { db200_hostpack(srcb + b0, b1 - b0, hc + b0 / 16, hv + b0 / 16); });
        g_dbg_wait_ms += std::chrono::duration<double, std::milli>(tw1 - tw0).count();
        g_dbg_pack_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw1).count();
        DB200_CUDA(cudaMemcpyAsync(pg->bases2.as<uint32_t>() + off / 16, hc, ngroups * 4, cudaMemcpyHostToDevice, up.cs2));
        DB200_CUDA(cudaMemcpyAsync(pg->nb.as<uint16_t>() + off / 16, hv, ngroups * 2, cudaMemcpyHostToDevice, up.cs2));
        DB200_CUDA(cudaEventRecord(up.pslot_done[sl], up.cs2));
        up.pslot_used[sl] = true;
        up.pslot_inflight[sl].store(true, std::memory_order_release);
        DB200_CUDA(cudaStreamWaitEvent(stream, up.pslot_done[sl], 0));
        return DB200_OK;
    };
    if (!host_pack) {
        for (uint64_t c = 0; c < nchunks; ++c) {
            DB200_TRY(enqueue_ascii(c, (int)(c & 1)));
            chunk_done[c] = 1;
            DB200_TRY(launch_ready_groups());
        }
    } else {
        // Two-ended schedule.  This thread packs from the front (32 Mbases per parallel pack) and enqueues the packed copies; a
        // FEEDER thread (page-locked sources only) watches the link and, whenever neither an ASCII copy nor a packed copy is in
        // flight, sends one 16 MB ASCII unit from the back — ASCII only fills the link time the packed copies leave idle, so the
        // split follows the measured host-pack and link speeds.  (Topping the ASCII queue up from this thread once per pack,
        // the previous version, kept the packed copies queued behind ASCII: 58 ms instead of 47 for 5 GB.)
        uint64_t front = 0, back = nchunks;
        std::mutex mu;                       // cursors, chunk_done / group_launched, launches of ready groups
        const bool dbg = std::getenv("DB200_DEBUG_UPLOAD") != nullptr;
        const auto t_begin = std::chrono::steady_clock::now();
        int feeder_rc = DB200_OK;
        std::string feeder_err;
        std::thread feeder;
        if (src_locked) feeder = std::thread([&] {
            cudaSetDevice(phys_of(device));
            for (;;) {
                { std::lock_guard<std::mutex> lk(mu); if (front >= back) return; }
                bool busy = false;
                for (int b2 = 0; b2 < Uploader::NSTAGE && !busy; ++b2)
                    if (stage_used[b2]) busy = cudaEventQuery(up.copied[b2]) == cudaErrorNotReady;
                for (int sl = 0; sl < Uploader::NPSLOT && !busy; ++sl)
                    if (up.pslot_inflight[sl].load(std::memory_order_acquire)) busy = cudaEventQuery(up.pslot_done[sl]) == cudaErrorNotReady;
                cudaGetLastError();
                // (sleeping, not spinning: every host CPU is packing; a spinning feeder slowed the pack threads by a third)
                if (busy) { std::this_thread::sleep_for(std::chrono::microseconds(30)); continue; }
                uint64_t c;
                { std::lock_guard<std::mutex> lk(mu); if (front >= back) return; c = --back; }
                int rc = enqueue_ascii(c, (int)(ascii_seq++ % Uploader::NSTAGE));
                if (rc == DB200_OK) { std::lock_guard<std::mutex> lk(mu); chunk_done[c] = 1; rc = launch_ready_groups(); }
                if (rc != DB200_OK) { feeder_rc = rc; feeder_err = t_err; std::lock_guard<std::mutex> lk(mu); back = front; return; }
            }
        });
        int rc_main = DB200_OK;
        for (;;) {
            uint64_t c, npk;
            { std::lock_guard<std::mutex> lk(mu); if (front >= back) break; npk = std::min<uint64_t>(2, back - front); c = front; front += npk; }
            rc_main = enqueue_packed(c, npk);
            if (rc_main == DB200_OK) { std::lock_guard<std::mutex> lk(mu); for (uint64_t u = c; u < c + npk; ++u) chunk_done[u] = 1; rc_main = launch_ready_groups(); }
            if (rc_main != DB200_OK) { std::lock_guard<std::mutex> lk(mu); back = front; break; }
        }
        if (feeder.joinable()) feeder.join();
        if (rc_main != DB200_OK) return rc_main;
        if (feeder_rc != DB200_OK) { set_error("%s", feeder_err.c_str()); return feeder_rc; }
        pg->ascii_chunks = ascii_seq; pg->host_packed_chunks = nchunks - ascii_seq;
        if (dbg) {
            const double t_all = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
            std::fprintf(stderr, "[db200 upload] %llu units of %llu bases: %llu host-packed, %llu ascii; enqueue loop %.2f ms (pack %.2f ms, slot waits %.2f ms), %u pack threads\n",
                         (unsigned long long)nchunks, (unsigned long long)CH, (unsigned long long)(nchunks - ascii_seq), (unsigned long long)ascii_seq, t_all, g_dbg_pack_ms, g_dbg_wait_ms,
                         up.packers ? up.packers->workers() + 1 : 0u);
            g_dbg_pack_ms = g_dbg_wait_ms = 0;
        }
    }
    DB200_TRY(launch_ready_groups());
    DB200_CUDA(cudaGetLastError());
    if (ps && up.ss) {
        const cudaError_t es = cudaStreamSynchronize(up.ss);
        if (es != cudaSuccess) { set_error("sketch (pipelined): %s", cudaGetErrorString(es)); return DB200_ECUDA; }
    }

    const cudaError_t e = cudaStreamSynchronize(stream);  // the host vectors above go out of scope
    if (e != cudaSuccess) { set_error("pack_genomes: %s", cudaGetErrorString(e)); return DB200_ECUDA; }
    return DB200_OK;
}

static int sketch_launch(const db200_packed_genomes *pg, int p, int canon, uint8_t *d_regs, cudaStream_t stream, uint32_t item_begin,
                         uint32_t item_count, int counter_idx) {
    const uint64_t m = 1ull << p;
    const int mode = (m * 4 <= (128u << 10)) ? 0 : (m <= (128u << 10)) ? 1 : 2;
    const size_t smem = mode == 0 ? m * 4 : mode == 1 ? m : 0;
    const int k = pg->k, kclass = k <= 16 ? 0 : (k < 32 ? 1 : 2), rcshift = 2 * (k - 1);
    SketchConsts kc;
    kc.four = 4u; kc.neg1 = 0xFFFFFFFFu;
    kc.rc_mul_lo = rcshift < 32 ? (1u << rcshift) : 0u;
    kc.rc_mul_hi = rcshift >= 32 ? (1u << (rcshift - 32)) : 0u;
    kc.rho_mul = 1u << p; kc.rho_add = 1u << (p - 1);
    kc.neg_2p20 = 0u - (1u << 20); kc.c1087_2p20 = 1087u << 20;
    int occ = 1;
    const unsigned sms = (unsigned)g_num_sms(pg->device);
#define DB200_SKETCH_LAUNCH(MODE, KC, CANON)                                                                                  \
    do {                                                                                                                      \
        auto kern = sketch_kernel<MODE, KC, CANON>;                                                                           \
        if (smem) DB200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 << 10));             \
        DB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SK_THREADS, smem));                              \
        const unsigned grid = (unsigned)std::min<uint64_t>(item_count, (uint64_t)sms * std::max(occ, 1));                     \
        kern<<<grid, SK_THREADS, smem, stream>>>(pg->bases2.as<uint4>(), pg->nb.as<uint64_t>(), pg->st.as<uint64_t>(),        \
                                                 pg->items.as<SketchItem>() + item_begin, item_count, pg->nblk - 1, k, p, d_regs, \
                                                 pg->counter.as<uint32_t>() + counter_idx, kc);                               \
    } while (0)
#define DB200_SKETCH_KC(MODE, CANON)                                        \
    do {                                                                    \
        if (kclass == 0) DB200_SKETCH_LAUNCH(MODE, 0, CANON);               \
        else if (kclass == 1) DB200_SKETCH_LAUNCH(MODE, 1, CANON);          \
        else DB200_SKETCH_LAUNCH(MODE, 2, CANON);                           \
    } while (0)
#define DB200_SKETCH_MODE(MODE)                                             \
    do {                                                                    \
        if (canon) DB200_SKETCH_KC(MODE, true);                             \
        else DB200_SKETCH_KC(MODE, false);                                  \
    } while (0)
    if (mode == 0) DB200_SKETCH_MODE(0);
    else if (mode == 1) DB200_SKETCH_MODE(1);
    else DB200_SKETCH_MODE(2);
#undef DB200_SKETCH_MODE
#undef DB200_SKETCH_KC
#undef DB200_SKETCH_LAUNCH
    DB200_LAUNCHED();
    DB200_CUDA(cudaGetLastError());
    return DB200_OK;
}

static int sketch_packed_impl(const db200_packed_genomes *pg, int p, int canon, uint8_t *d_regs, cudaStream_t stream) {
    if (p < 7 || p > 24) { set_error("sketch: p=%d outside the GPU path's range [7,24]", p); return DB200_EUNSUPPORTED; }
    if (pg->k < 1 || pg->k > 32) { set_error("sketch: k=%d outside [1,32]", pg->k); return DB200_EUNSUPPORTED; }
    DB200_CUDA(cudaMemsetAsync(d_regs, 0, pg->ngenomes << p, stream));
    if (pg->nitems == 0) return DB200_OK;
    const int ci = 8 + (int)(pg->launch_seq.fetch_add(1) % 8u);
    DB200_CUDA(cudaMemsetAsync(pg->counter.as<uint32_t>() + ci, 0, 4, stream));
    return sketch_launch(pg, p, canon, d_regs, stream, 0, pg->nitems, ci);
}

// ---------------------------------------------------------------------------------------------
// TMA descriptor
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

} // namespace db200

// ---------------------------------------------------------------------------------------------
// dist plan
// ---------------------------------------------------------------------------------------------
struct db200_dist_plan {
    int device = 0;
    uint64_t nrows = 0;            // rows of the plane tensor (symmetric: n; rect: qbase + nq)
    uint64_t n1 = 0, qbase = 0, n2 = 0;  // valid row segments [0,n1) and [qbase, qbase+n2)
    int p = 0, estim = -1, gmin = 0, gmax = 0, K = 0;
    bool ready = false;
    CUtensorMap tmap, tmap16;   // boxes of 32 / 16 sketches
    static constexpr int NSLOT = 4;   // independent tile lists, so that row blocks of one request can be in flight together
    db200::DevBuf planes, counts, card, smin, smax, pmin, pmax, minmax, lists, sthr, pthr, llists, ptl;
    db200::DevBuf tiles[NSLOT];
    db200::DevBuf knn_vals, knn_keys;   // k-NN: one row block of all-pairs values, retained (value, index) keys
    std::vector<db200::DistTile> host_tiles[NSLOT];
    // tile-list cache key
    struct TileKey { int rect = -1, ta = 0; uint64_t rb = 0, re = 0, nr = 0, nq = 0, n = 0, ntiles = 0; } tl[NSLOT];
    uint64_t last_pairs = 0, last_tiles = 0;
    std::mutex mu;
};

namespace db200 {

// A plan is built in three steps so that the multi-process driver can overlap the plane build with the exchange of the register
// shards: begin (allocations for a KNOWN global register range), add_rows (planes + counts + tails + cardinalities of a row
// range, as soon as those rows are on the device), finish (tensor maps).  plan_prepare is the one-shot form.
static int plan_begin(db200_dist_plan *pl, uint64_t nrows, uint64_t n1, uint64_t qbase, uint64_t n2, int p, int estim, int gmin, int gmax,
                      cudaStream_t stream) {
    if (p < 7 || p > 20) { set_error("dist: p=%d outside the GPU path's range [7,20]", p); return DB200_EUNSUPPORTED; }
    if (estim < 0 || estim > 2) { set_error("dist: unknown estimation method %d", estim); return DB200_EINVAL; }
    if (nrows == 0 || nrows > (1ull << 31) - 64) { set_error("dist: %llu sketches unsupported", (unsigned long long)nrows); return DB200_EINVAL; }
    if (gmin < 0 || gmax < gmin) { set_error("dist: bad register range [%d,%d]", gmin, gmax); return DB200_EINVAL; }
    if (gmax > 64 - p + 1) { set_error("dist: register value %d exceeds 64-p+1=%d (corrupt sketch?)", gmax, 64 - p + 1); return DB200_EINVAL; }
    if (!get_encode()) { set_error("cuTensorMapEncodeTiled not available from the driver"); return DB200_ECUDA; }
    pl->ready = false;
    pl->nrows = nrows; pl->n1 = n1; pl->qbase = qbase; pl->n2 = n2; pl->p = p; pl->estim = estim;
    // tile lists depend only on (n, row range, mode), not on the register data: they stay cached across prepares
    const uint64_t m = 1ull << p, W = std::max<uint64_t>(m >> 5, 32), npan = (nrows + DT - 1) / DT;
    pl->gmin = gmin; pl->gmax = gmax; pl->K = gmax - gmin;
    const uint64_t Kalloc = std::max(pl->K, 1);
    DB200_TRY(pl->planes.reserve(Kalloc * nrows * W * 4));
    DB200_TRY(pl->counts.reserve(nrows * 64 * 4));
    DB200_TRY(pl->card.reserve(nrows * 8));
    DB200_TRY(pl->smin.reserve(nrows));
    DB200_TRY(pl->smax.reserve(nrows));
    DB200_TRY(pl->pmin.reserve(npan * 4));
    DB200_TRY(pl->pmax.reserve(npan * 4));
    DB200_TRY(pl->pthr.reserve(npan * 4));
    DB200_TRY(pl->lists.reserve(nrows * SPARSE_C * 4));
    DB200_TRY(pl->sthr.reserve(nrows));
    DB200_CUDA(cudaMemsetAsync(pl->pthr.ptr, 0, npan * 4, stream));
    DB200_TRY(pl->ptl.reserve(npan * 4));
    DB200_TRY(pl->llists.reserve(nrows * SPARSE_C * 4));
    DB200_CUDA(cudaMemsetAsync(pl->ptl.ptr, 0xFF, npan * 4, stream));
    DB200_CUDA(cudaMemsetAsync(pl->pmin.ptr, 0xFF, npan * 4, stream));
    DB200_CUDA(cudaMemsetAsync(pl->pmax.ptr, 0, npan * 4, stream));
    if (n1 + n2 != nrows || pl->K == 0) {  // padding rows (or the K==0 dummy plane) must read as "below every threshold"
        DB200_CUDA(cudaMemsetAsync(pl->planes.ptr, 0, Kalloc * nrows * W * 4, stream));
        DB200_CUDA(cudaMemsetAsync(pl->smin.ptr, 0, nrows, stream));
        DB200_CUDA(cudaMemsetAsync(pl->smax.ptr, 0, nrows, stream));
        DB200_CUDA(cudaMemsetAsync(pl->card.ptr, 0, nrows * 8, stream));
        DB200_CUDA(cudaMemsetAsync(pl->lists.ptr, 0, nrows * SPARSE_C * 4, stream));
        DB200_CUDA(cudaMemsetAsync(pl->llists.ptr, 0, nrows * SPARSE_C * 4, stream));
    }
    DB200_CUDA(cudaMemsetAsync(pl->counts.ptr, 0, nrows * 64 * 4, stream));
    return DB200_OK;
}

// rows [r0, r0 + cnt) of the register matrix at d_regs (row-major base of ALL nrows rows): planes, counts, tails, cardinalities
static int plan_add_rows(db200_dist_plan *pl, const uint8_t *d_regs, uint64_t r0, uint64_t cnt, cudaStream_t stream) {
    if (!cnt) return DB200_OK;
    if (r0 + cnt > pl->nrows) { set_error("dist plan: rows [%llu,%llu) outside the plan's %llu", (unsigned long long)r0, (unsigned long long)(r0 + cnt), (unsigned long long)pl->nrows); return DB200_EINVAL; }
    if (pl->K > 0) {
        planes_kernel<<<(unsigned)cnt, 128, 0, stream>>>(reinterpret_cast<const uint32_t *>(d_regs), pl->nrows, r0, pl->p, pl->gmin, pl->K,
                                                        pl->planes.as<uint32_t>(), pl->counts.as<uint32_t>(), pl->lists.as<uint32_t>(),
                                                        pl->sthr.as<uint8_t>(), pl->pthr.as<uint32_t>(), pl->llists.as<uint32_t>(),
                                                        pl->ptl.as<uint32_t>());
        DB200_LAUNCHED();
    }
    card_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, stream>>>(pl->counts.as<uint32_t>(), r0, cnt, pl->p, pl->gmin, pl->gmax, pl->estim,
                                                                  pl->card.as<double>(), pl->smin.as<uint8_t>(), pl->smax.as<uint8_t>(),
                                                                  pl->pmin.as<uint32_t>(), pl->pmax.as<uint32_t>());
    DB200_LAUNCHED();
    DB200_CUDA(cudaGetLastError());
    return DB200_OK;
}

static int plan_finish(db200_dist_plan *pl) {
    PFN_encodeTiled enc = get_encode();
    const uint64_t m = 1ull << pl->p, W = std::max<uint64_t>(m >> 5, 32), Kalloc = std::max(pl->K, 1);
    // 3-D tensor {W words, nrows sketches, K thresholds}, box {32, 32, 1}, 128-byte swizzle
    cuuint64_t dims[3] = {W, pl->nrows, Kalloc};
    cuuint64_t strides[2] = {W * 4, pl->nrows * W * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)DT, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&pl->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, pl->planes.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return DB200_ECUDA; }
    box[1] = (cuuint32_t)JT;
    r = enc(&pl->tmap16, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, pl->planes.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (16-row box) failed with CUresult %d", (int)r); return DB200_ECUDA; }
    pl->ready = true;
    return DB200_OK;
}

static int plan_prepare(db200_dist_plan *pl, const uint8_t *d_regs, uint64_t nrows, uint64_t n1, uint64_t qbase, uint64_t n2, int p,
                        int estim, cudaStream_t stream) {
    if (p < 7 || p > 20) { set_error("dist: p=%d outside the GPU path's range [7,20]", p); return DB200_EUNSUPPORTED; }
    pl->ready = false;
    const uint64_t m = 1ull << p;
    DB200_TRY(pl->minmax.reserve(8));
    DB200_CUDA(cudaMemsetAsync(pl->minmax.ptr, 0xFF, 4, stream));
    DB200_CUDA(cudaMemsetAsync(pl->minmax.as<uint8_t>() + 4, 0, 4, stream));
    const int sms = g_num_sms(pl->device);
    for (int seg = 0; seg < 2; ++seg) {
        const uint64_t r0 = seg ? qbase : 0, cnt = seg ? n2 : n1;
        if (!cnt) continue;
        const uint64_t n16 = cnt * m / 16;
        const unsigned grid = (unsigned)std::min<uint64_t>((n16 + 255) / 256, (uint64_t)sms * 8);
        range_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(d_regs + r0 * m), n16, pl->minmax.as<uint32_t>());
        DB200_LAUNCHED();
    }
    uint32_t mm[2];
    DB200_CUDA(cudaMemcpyAsync(mm, pl->minmax.ptr, 8, cudaMemcpyDeviceToHost, stream));
    DB200_CUDA(cudaStreamSynchronize(stream));
    if (mm[1] > (uint32_t)(64 - p + 1)) { set_error("dist: register value %u exceeds 64-p+1=%d (corrupt sketch?)", mm[1], 64 - p + 1); return DB200_EINVAL; }
    if (mm[0] > mm[1]) mm[0] = mm[1] = 0;      // (no rows at all)
    DB200_TRY(plan_begin(pl, nrows, n1, qbase, n2, p, estim, (int)mm[0], (int)mm[1], stream));
    DB200_TRY(plan_add_rows(pl, d_regs, 0, n1, stream));
    DB200_TRY(plan_add_rows(pl, d_regs, qbase, n2, stream));
    return plan_finish(pl);
}

static int plan_tiles(db200_dist_plan *pl, int slot, int rect, int ta, uint64_t rb, uint64_t re, uint64_t nr, uint64_t nq, cudaStream_t stream) {
    auto &key = pl->tl[slot];
    if (key.rect == rect && key.ta == ta && key.rb == rb && key.re == re && key.nr == nr && key.nq == nq && key.n == pl->nrows) return DB200_OK;
    // A panels have `ta` rows (32, or 16 for the joint-MLE kernel), B panels DT = 32.  Tiles are ordered in
    // super-rows of A panels: concurrent CTAs share their A panels and sweep the B panels together (L2 reuse).
    std::vector<DistTile> &tiles = pl->host_tiles[slot];
    tiles.clear();
    const uint32_t f = (uint32_t)(DT / ta);          // A panels per B panel
    const uint32_t SR = 8 * f;
    if (!rect) {
        const uint64_t n = pl->nrows;
        const uint32_t nb = (uint32_t)((n + DT - 1) / DT);
        const uint32_t a0 = (uint32_t)(rb / ta), a1 = re > rb ? (uint32_t)((re - 1) / ta) + 1 : a0;
        for (uint32_t sr = a0; sr < a1; sr += SR)
            for (uint32_t b = sr / f; b < nb; ++b)
                for (uint32_t a = sr; a < std::min(sr + SR, a1) && a / f <= b; ++a) tiles.push_back(DistTile{a, b});
    } else {
        const uint32_t na = (uint32_t)((nq + ta - 1) / ta), nb = (uint32_t)((nr + DT - 1) / DT);
        for (uint32_t sr = 0; sr < na; sr += SR)
            for (uint32_t b = 0; b < nb; ++b)
                for (uint32_t a = sr; a < std::min(sr + SR, na); ++a) tiles.push_back(DistTile{a, b});
    }
    key.ntiles = tiles.size();
    if (!tiles.empty()) {
        DB200_TRY(pl->tiles[slot].reserve(tiles.size() * sizeof(DistTile)));
        // pageable source: the runtime stages it before returning; host_tiles[slot] also stays alive in the plan
        DB200_CUDA(cudaMemcpyAsync(pl->tiles[slot].ptr, tiles.data(), tiles.size() * sizeof(DistTile), cudaMemcpyHostToDevice, stream));
    }
    key.rect = rect; key.ta = ta; key.rb = rb; key.re = re; key.nr = nr; key.nq = nq; key.n = pl->nrows;
    return DB200_OK;
}

static int plan_run(db200_dist_plan *pl, const db200_dist_params *prm, int rect, uint64_t rb, uint64_t re, uint64_t nr, uint64_t nq,
                    float *d_out, cudaStream_t stream, int slot = 0, bool ksinv_double = false) {
    if (!pl->ready) { set_error("dist plan not prepared"); return DB200_EINVAL; }
    if (prm->p != pl->p || prm->estim != pl->estim) { set_error("dist params (p=%d, estim=%d) differ from the prepared plan (p=%d, estim=%d)", prm->p, prm->estim, pl->p, pl->estim); return DB200_EINVAL; }
    if (prm->result_type < 0 || prm->result_type > DB200_UNION_SIZE) { set_error("dist: unknown result type %d", prm->result_type); return DB200_EINVAL; }
    if (prm->k < 1) { set_error("dist: k must be positive"); return DB200_EINVAL; }
    const bool joint = prm->jestim == DB200_ERTL_JOINT_MLE;
    DB200_TRY(plan_tiles(pl, slot, rect, joint ? JT : DT, rb, re, nr, nq, stream));
    const uint64_t ntiles = pl->tl[slot].ntiles;
    pl->last_tiles = ntiles;
    if (ntiles == 0) { pl->last_pairs = 0; return DB200_OK; }
    DistArgs a;
    a.tiles = pl->tiles[slot].as<DistTile>();
    a.smin = pl->smin.as<uint8_t>(); a.smax = pl->smax.as<uint8_t>();
    a.pmin = pl->pmin.as<uint32_t>(); a.pmax = pl->pmax.as<uint32_t>();
    a.card = pl->card.as<double>();
    a.counts = pl->counts.as<uint32_t>(); a.lists = pl->lists.as<uint32_t>(); a.pthr = pl->pthr.as<uint32_t>();
    a.llists = pl->llists.as<uint32_t>(); a.ptl = pl->ptl.as<uint32_t>(); a.low = 0;
    a.out = d_out;
    a.n = pl->nrows; a.row_begin = rb; a.row_end = re;
    a.out_base = rect ? 0 : (rb * (2 * pl->nrows - rb - 1)) / 2;
    a.nr = nr; a.nq = nq; a.qbase = pl->qbase;
    // const float ksinv = 1./k in dist_loop (src/sketch_and_cmp.h:797) but const double in nndist_loop (:729)
    a.ksinv = ksinv_double ? 1. / prm->k : (double)(float)(1. / prm->k);
    a.p = pl->p; a.gmin = pl->gmin; a.gmax = pl->gmax; a.K = pl->K;
    a.estim = prm->estim; a.rtype = prm->result_type; a.rect = rect; a.one = 1;
    const bool wide = pl->p > 16;                  // threshold counts above 2^16: 32-bit count storage
    const size_t gsz = wide ? 4 : 2;
    if (!joint) {
        // shared memory: S stages of 8 KiB + (K + 1) x 1024 counts + barriers; aim for two CTAs per SM
        const size_t gbytes = (size_t)(pl->K + 1) * DT * DT * gsz;   // bins lo..hi of a tile: at most K + 1
        int S = 6;
        const size_t budget2 = 113 << 10, budget1 = 226 << 10;
        if (gbytes + 2 * STAGE_BYTES + 1024 > budget1) { set_error("dist: %d live thresholds do not fit in shared memory at p=%d", pl->K, pl->p); return DB200_EUNSUPPORTED; }
        // two CTAs per SM matter more than pipeline depth: drop to 5 or 4 stages before giving up the second CTA
        while (S > 4 && gbytes + (size_t)S * STAGE_BYTES + 1024 > budget2) --S;
        if (gbytes + (size_t)S * STAGE_BYTES + 1024 > budget2) S = (int)std::min<size_t>(12, (budget1 - gbytes - 1024) / STAGE_BYTES);
        a.stages = S;
        a.sparse = S >= 4;   // the stage buffers double as sparse-tail storage (16 + 3.6 KB); with fewer stages every threshold is dense
        a.low = S >= 5;      // ... plus 16 KB for the low tails
        const size_t smem = (size_t)S * STAGE_BYTES + gbytes + 2 * S * 8;
        if (wide) {
            DB200_CUDA(cudaFuncSetAttribute(dist_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 << 10));
            dist_kernel<uint32_t><<<(unsigned)ntiles, DIST_THREADS, smem, stream>>>(pl->tmap, a);
        } else {
            DB200_CUDA(cudaFuncSetAttribute(dist_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 << 10));
            dist_kernel<uint16_t><<<(unsigned)ntiles, DIST_THREADS, smem, stream>>>(pl->tmap, a);
        }
    } else {
        const size_t gbytes = (size_t)3 * std::max(pl->K, 1) * JPAIRS * gsz;
        const size_t budget1 = 226 << 10;
        if (gbytes + 2 * JSTAGE_BYTES + 1024 > budget1) { set_error("dist (joint MLE): %d live thresholds do not fit in shared memory at p=%d", pl->K, pl->p); return DB200_EUNSUPPORTED; }
        const int S = (int)std::min<size_t>(6, (budget1 - gbytes - 1024) / JSTAGE_BYTES);   // stage buffers double as sparse-tail storage (22 KB)
        a.stages = S;
        a.sparse = 1;      // 2 joint stages (24 KB) already hold the 22 KB of staged tails
        const size_t smem = (size_t)S * JSTAGE_BYTES + gbytes + 2 * S * 8;
        const int lhs_is_b = rect ? 1 : (prm->order == DB200_ORDER_COL_FIRST ? 1 : 0);
        if (wide) {
            DB200_CUDA(cudaFuncSetAttribute(dist_jmle_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 << 10));
            dist_jmle_kernel<uint32_t><<<(unsigned)ntiles, JTHREADS, smem, stream>>>(pl->tmap16, pl->tmap, a, lhs_is_b);
        } else {
            DB200_CUDA(cudaFuncSetAttribute(dist_jmle_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 << 10));
            dist_jmle_kernel<uint16_t><<<(unsigned)ntiles, JTHREADS, smem, stream>>>(pl->tmap16, pl->tmap, a, lhs_is_b);
        }
    }
    DB200_LAUNCHED();
    DB200_CUDA(cudaGetLastError());
    if (rect) pl->last_pairs = nr * nq;
    else {
        const uint64_t n = pl->nrows;
        auto tri = [n](uint64_t r) { return (r * (2 * n - r - 1)) / 2; };
        pl->last_pairs = tri(std::min(re, n)) - tri(rb);
    }
    return DB200_OK;
}

// emt2nntype, src/dashing.h:268-280
static bool is_similarity(int rtype) {
    switch (rtype) {
        case DB200_MASH_DIST: case DB200_FULL_MASH_DIST: case DB200_CONTAINMENT_DIST: case DB200_FULL_CONTAINMENT_DIST:
        case DB200_SYMMETRIC_CONTAINMENT_DIST: return false;
        default: return true;
    }
}

// k nearest neighbours: the all-pairs values of one row block at a time go to plan scratch, a second kernel folds them
// into the per-sketch retained sets, a third sorts and decodes.  nq == 0: symmetric.
static int plan_knn(db200_dist_plan *pl, const db200_dist_params *prm, uint64_t nr, uint64_t nq, uint32_t nn, Neighbor *d_out,
                    cudaStream_t stream, uint64_t sym_row_begin = 0, uint64_t sym_row_end = ~0ull) {
    if (!pl->ready) { set_error("dist plan not prepared"); return DB200_EINVAL; }
    if (nn < 1 || nn > 1024) { set_error("nearest neighbours: nneighbors=%u outside the GPU path's range [1,1024]", nn); return DB200_EUNSUPPORTED; }
    if (prm->result_type < 0 || prm->result_type > DB200_UNION_SIZE) { set_error("dist: unknown result type %d", prm->result_type); return DB200_EINVAL; }
    const int rect = nq != 0, sim = is_similarity(prm->result_type);
    const uint64_t n = pl->nrows, rows = rect ? nq : n;
    // all-pairs values held in HBM at a time (1 GiB of floats unless DB200_KNN_BLOCK_PAIRS says otherwise)
    const char *benv = std::getenv("DB200_KNN_BLOCK_PAIRS");
    const uint64_t bval = benv ? std::strtoull(benv, nullptr, 10) : 0;
    const uint64_t budget = bval ? bval : (uint64_t)1 << 28;
    DB200_TRY(pl->knn_keys.reserve(rows * nn * 8));
    knn_init_kernel<<<(unsigned)std::min<uint64_t>((rows * nn + 255) / 256, 65535), 256, 0, stream>>>(pl->knn_keys.as<uint64_t>(), rows * nn, sim);
    DB200_LAUNCHED();
    KnnArgs ka{};
    ka.keys = pl->knn_keys.as<uint64_t>(); ka.n = n; ka.nr = nr; ka.nn = nn; ka.sim = sim; ka.rect = rect;
    const size_t smem = (size_t)KNN_WARPS * nn * 8;
    uint64_t total_pairs = 0, total_tiles = 0;
    if (!rect) {
        auto tri = [n](uint64_t r) { return (r * (2 * n - r - 1)) / 2; };
        // (a row range: the partial table of the pairs (i, j > i) with i in the range — what one rank of the multi-process
        // driver contributes; the full table is the range [0, n))
        const uint64_t r_end = std::min(n, sym_row_end);
        for (uint64_t rb = std::min(sym_row_begin, r_end); rb + 1 < n && rb < r_end;) {
            // as many whole panels of rows as fit the budget (at least one)
            uint64_t re = std::min(r_end, (rb / DT + 1) * DT);
            while (re < r_end && tri(std::min(r_end, re + DT)) - tri(rb) <= budget) re = std::min(r_end, re + DT);
            const uint64_t npairs = tri(re) - tri(rb);
            if (npairs) {
                DB200_TRY(pl->knn_vals.reserve(npairs * 4));
                DB200_TRY(plan_run(pl, prm, 0, rb, re, 0, 0, pl->knn_vals.as<float>(), stream, 0, true));
                total_pairs += pl->last_pairs; total_tiles += pl->last_tiles;
                ka.vals = pl->knn_vals.as<float>(); ka.rb = rb; ka.re = re;
                knn_update_kernel<<<(unsigned)((n - rb + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, smem, stream>>>(ka);
                DB200_LAUNCHED();
            }
            rb = re;
        }
    } else {
        if (nr == 0) { set_error("nearest neighbours: no references"); return DB200_EINVAL; }
        const uint64_t qbase0 = pl->qbase;
        uint64_t qstep = std::max<uint64_t>(DT, budget / nr / DT * DT);
        int rc = DB200_OK;
        for (uint64_t q0 = 0; q0 < nq && rc == DB200_OK; q0 += qstep) {
            const uint64_t qn = std::min(qstep, nq - q0);
            rc = pl->knn_vals.reserve(qn * nr * 4);
            if (rc != DB200_OK) break;
            pl->qbase = qbase0 + q0;     // queries [q0, q0 + qn) of the plan as a rect run of their own (q0 % DT == 0)
            rc = plan_run(pl, prm, 1, 0, 0, nr, qn, pl->knn_vals.as<float>(), stream, 0, true);
            pl->qbase = qbase0;
            if (rc != DB200_OK) break;
            total_pairs += pl->last_pairs; total_tiles += pl->last_tiles;
            ka.vals = pl->knn_vals.as<float>(); ka.q0 = q0; ka.nq = qn;
            knn_update_kernel<<<(unsigned)((qn + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, smem, stream>>>(ka);
            DB200_LAUNCHED();
        }
        if (rc != DB200_OK) return rc;
    }
    knn_sort_kernel<<<(unsigned)((rows + KNN_WARPS - 1) / KNN_WARPS), KNN_WARPS * 32, smem, stream>>>(pl->knn_keys.as<uint64_t>(), rows, nn, sim, d_out);
    DB200_LAUNCHED();
    DB200_CUDA(cudaGetLastError());
    pl->last_pairs = total_pairs; pl->last_tiles = total_tiles;
    return DB200_OK;
}

// Caller-supplied per-sketch cardinalities replace the ones card_kernel evaluated (the reference's hll_t objects carry a cached
// value_: sketches loaded from files keep what read() computed under the FILE's estimator, hll.h:1078).  Rows [0, n1) and
// [qbase, qbase + n2) of the plan.  Pageable sources are staged by the runtime before cudaMemcpyAsync returns.
static int plan_override_cards(db200_dist_plan *pl, const double *c1, uint64_t n1, const double *c2, uint64_t n2, cudaStream_t stream) {
    if (c1 && n1) DB200_CUDA(cudaMemcpyAsync(pl->card.as<double>(), c1, n1 * 8, cudaMemcpyHostToDevice, stream));
    if (c2 && n2) DB200_CUDA(cudaMemcpyAsync(pl->card.as<double>() + pl->qbase, c2, n2 * 8, cudaMemcpyHostToDevice, stream));
    return DB200_OK;
}

// Default per-device resources for the host-pointer entry points.
struct HostCtx {
    std::mutex mu;
    cudaStream_t stream = nullptr, cstream = nullptr;
    cudaEvent_t blk_done[db200_dist_plan::NSLOT] = {nullptr, nullptr, nullptr, nullptr};
    DevBuf regs, out, cards, nbrs;
    static constexpr int NSTREAM = 3;          // row-block streaming: blocks in flight (kernel / copy / callback)
    char *stream_host[NSTREAM] = {nullptr, nullptr, nullptr};
    size_t stream_cap = 0;
    cudaEvent_t stream_copied[NSTREAM] = {nullptr, nullptr, nullptr};
    DevBuf fa_text, fa_sums, fa_wsum, fa_state, fa_pos, fa_tab, fa_gend, fa_carry, fa_flags;   // device-side FASTA parsing (fasta.cuh)
    Uploader up;
    std::unique_ptr<db200_dist_plan> plan;
    std::unique_ptr<db200_packed_genomes> store;   // reused by db200_sketch_batch
    int init(int device) {
        if (!stream) DB200_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        if (!cstream) {
            DB200_CUDA(cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking));
            for (auto &e : blk_done) DB200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        if (!plan) { plan.reset(new db200_dist_plan); plan->device = device; }
        if (!store) store.reset(new db200_packed_genomes);
        return DB200_OK;
    }
};
static HostCtx &host_ctx(int device) {
    // leaked on purpose: static destructors would call cudaFree / cudaStreamDestroy after the CUDA runtime may already have
    // been torn down at process exit (undefined behaviour).  check_device() has bounded `device` by logical_device_count() <= 64.
    static HostCtx *ctx = new HostCtx[64];
    return ctx[device & 63];
}

// device = DB200_ALL_DEVICES: one host thread per logical device, each driving the single-device entry point on its share
// of the units (genomes, sketches, block rows, queries).  The host already holds every input, so no device-to-device
// exchange is needed; the first failure (with its device number) becomes the caller's error.
static int for_each_device(int ndev, const std::function<int(int)> &fn) {
    if (ndev <= 0) return check_device(0);
    std::vector<int> rc((size_t)ndev, DB200_OK);
    std::vector<std::string> err((size_t)ndev);
    std::vector<std::thread> th;
    for (int d = 1; d < ndev; ++d)
        th.emplace_back([&, d] { rc[d] = fn(d); if (rc[d] != DB200_OK) err[d] = t_err; });
    rc[0] = fn(0);
    if (rc[0] != DB200_OK) err[0] = t_err;
    for (auto &t : th) t.join();
    for (int d = 0; d < ndev; ++d)
        if (rc[d] != DB200_OK) { set_error("device %d: %s", d, err[d].c_str()); return rc[d]; }
    return DB200_OK;
}

// contiguous split of [0, n) into `parts` ranges of (nearly) equal weight; weight(i) = prefix weight of the first i units
static std::vector<uint64_t> split_by_weight(uint64_t n, int parts, const std::function<uint64_t(uint64_t)> &prefix) {
    std::vector<uint64_t> cut((size_t)parts + 1, n);
    cut[0] = 0;
    const uint64_t total = prefix(n);
    uint64_t at = 0;
    for (int d = 1; d < parts; ++d) {
        const uint64_t target = total / (uint64_t)parts * (uint64_t)d + total % (uint64_t)parts * (uint64_t)d / (uint64_t)parts;
        uint64_t lo = at, hi = n;           // first i with prefix(i) >= target
        while (lo < hi) { const uint64_t mid = lo + (hi - lo) / 2; if (prefix(mid) >= target) hi = mid; else lo = mid + 1; }
        cut[d] = at = lo;
    }
    return cut;
}

} // namespace db200

using namespace db200;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char *db200_last_error(void) { return t_err.c_str(); }
int db200_version(void) { return DB200_VERSION; }
uint64_t db200_kernel_launches(void) { return g_kernel_launches.load(); }

int db200_device_count(void) {
    return logical_device_count();
}

int db200_warmup(int device) {
    if (device == DB200_ALL_DEVICES) {
        return for_each_device(logical_device_count(), [](int d) { return db200_warmup(d); });
    }
    DB200_TRY(check_device(device));
    DB200_CUDA(cudaFree(nullptr));
    return DB200_OK;
}

int db200_host_alloc(void **out, size_t bytes) {
    if (!out) { set_error("db200_host_alloc: null out"); return DB200_EINVAL; }
    // On the calling thread's CURRENT device, which is left alone: a rank of a multi-process job that had selected GPU r found
    // device 0 current after this call (round 2, 8-GPU run: the host's next CUDA event then belonged to the wrong device).
    // Page-locked memory is usable from every device under unified addressing.
    if (phys_device_count() == 0) { set_error("no usable CUDA device; libdashing_b200 has no CPU fallback"); return DB200_ENODEV; }
    DB200_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return DB200_OK;
}
int db200_host_free(void *ptr) {
    if (ptr) DB200_CUDA(cudaFreeHost(ptr));
    return DB200_OK;
}

// ---- sketching ------------------------------------------------------------------------------
int db200_pack_genomes(int device, const char *bases, const uint64_t *rec_offsets, uint64_t nrecords, const uint64_t *genome_rec_begin,
                       uint64_t ngenomes, int k, db200_packed_genomes **out) {
    if (!out || !rec_offsets || !genome_rec_begin || (!bases && nrecords)) { set_error("db200_pack_genomes: null argument"); return DB200_EINVAL; }
    if (k < 1 || k > 32) { set_error("sketch: k=%d outside [1,32]", k); return DB200_EUNSUPPORTED; }
    if (ngenomes >= (1ull << 32)) { set_error("too many genomes"); return DB200_EINVAL; }
    DB200_TRY(check_device(device));
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    std::unique_ptr<db200_packed_genomes> pg(new db200_packed_genomes);
    DB200_TRY(pack_genomes_impl(device, bases, rec_offsets, nrecords, genome_rec_begin, ngenomes, k, pg.get(), hc.up, hc.stream));
    *out = pg.release();
    return DB200_OK;
}

int db200_repack_genomes(db200_packed_genomes *g, const char *bases, const uint64_t *rec_offsets, uint64_t nrecords,
                         const uint64_t *genome_rec_begin, uint64_t ngenomes, int k) {
    if (!g || !rec_offsets || !genome_rec_begin || (!bases && nrecords)) { set_error("db200_repack_genomes: null argument"); return DB200_EINVAL; }
    if (k < 1 || k > 32) { set_error("sketch: k=%d outside [1,32]", k); return DB200_EUNSUPPORTED; }
    if (ngenomes >= (1ull << 32)) { set_error("too many genomes"); return DB200_EINVAL; }
    const int device = g->device;
    DB200_TRY(check_device(device));
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    return pack_genomes_impl(device, bases, rec_offsets, nrecords, genome_rec_begin, ngenomes, k, g, hc.up, hc.stream);
}

int db200_packed_genomes_free(db200_packed_genomes *g) {
    if (g) { cudaSetDevice(phys_of(g->device)); delete g; }
    return DB200_OK;
}

int db200_packed_genomes_stats(const db200_packed_genomes *g, uint64_t *packed_bytes, uint64_t *kmers, uint64_t *bases) {
    if (!g) { set_error("null store"); return DB200_EINVAL; }
    // algorithmic bytes of SURVEY.md §8(d): ceil(L/4) 2-bit bases + ceil(L/8) validity plane (+ the record-start plane this design adds)
    if (packed_bytes) *packed_bytes = (g->nbases + 3) / 4 + 2 * ((g->nbases + 7) / 8);
    if (kmers) *kmers = g->kmers;
    if (bases) *bases = g->nbases;
    return DB200_OK;
}

int db200_sketch_packed_dev(const db200_packed_genomes *g, int p, int canon, uint8_t *d_registers, void *stream) {
    if (!g || !d_registers) { set_error("db200_sketch_packed_dev: null argument"); return DB200_EINVAL; }
    DB200_TRY(check_device(g->device));
    return sketch_packed_impl(g, p, canon, d_registers, (cudaStream_t)stream);
}

static int sketch_batch_all(int p, int k, int canon, const char *bases, const uint64_t *rec_offsets, uint64_t nrecords,
                            const uint64_t *genome_rec_begin, uint64_t ngenomes, uint8_t *registers_out);

int db200_sketch_batch(int device, int p, int k, int canon, const char *bases, const uint64_t *rec_offsets, uint64_t nrecords,
                       const uint64_t *genome_rec_begin, uint64_t ngenomes, uint8_t *registers_out) {
    if (device == DB200_ALL_DEVICES && logical_device_count() > 1 && rec_offsets && genome_rec_begin && ngenomes > 1)
        return sketch_batch_all(p, k, canon, bases, rec_offsets, nrecords, genome_rec_begin, ngenomes, registers_out);
    if (device == DB200_ALL_DEVICES) device = 0;
    if (!registers_out && ngenomes) { set_error("db200_sketch_batch: null output"); return DB200_EINVAL; }
    if (!rec_offsets || !genome_rec_begin || (!bases && nrecords)) { set_error("db200_sketch_batch: null argument"); return DB200_EINVAL; }
    if (k < 1 || k > 32) { set_error("sketch: k=%d outside [1,32]", k); return DB200_EUNSUPPORTED; }
    if (p < 7 || p > 24) { set_error("sketch: p=%d outside the GPU path's range [7,24]", p); return DB200_EUNSUPPORTED; }
    if (ngenomes >= (1ull << 32)) { set_error("too many genomes"); return DB200_EINVAL; }
    DB200_TRY(check_device(device));
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    db200_packed_genomes *pg = hc.store.get();   // device buffers persist across calls
    const uint64_t bytes = ngenomes << p;
    DB200_TRY(hc.regs.reserve(std::max<uint64_t>(bytes, 16)));
    DB200_TRY(hc.up.init());
    const PipelinedSketch ps{p, canon, hc.regs.as<uint8_t>()};   // each genome group is sketched while the next ones upload
    DB200_TRY(pack_genomes_impl(device, bases, rec_offsets, nrecords, genome_rec_begin, ngenomes, k, pg, hc.up, hc.stream, &ps));
    DB200_CUDA(cudaMemcpyAsync(registers_out, hc.regs.ptr, bytes, cudaMemcpyDeviceToHost, hc.stream));
    DB200_CUDA(cudaStreamSynchronize(hc.stream));
    return DB200_OK;
}

} // extern "C"

// streaming sketcher: per-slot host staging of the records handed over by the reference's kseq loop
struct db200_sketcher {
    int p, k, canon, device;
    struct Slot { std::vector<char> bases; std::vector<uint64_t> offs{0}; };
    std::vector<Slot> slots;
};

extern "C" {

int db200_sketcher_create(int p, int k, int canon, int device, uint32_t nslots, db200_sketcher **out) {
    if (!out || nslots == 0) { set_error("db200_sketcher_create: bad arguments"); return DB200_EINVAL; }
    if (k < 1 || k > 32) { set_error("sketch: k=%d outside [1,32]", k); return DB200_EUNSUPPORTED; }
    if (p < 7 || p > 24) { set_error("sketch: p=%d outside the GPU path's range [7,24]", p); return DB200_EUNSUPPORTED; }
    DB200_TRY(check_device(device == DB200_ALL_DEVICES ? 0 : device));
    db200_sketcher *h = new db200_sketcher{p, k, canon, device, {}};
    h->slots.resize(nslots);
    *out = h;
    return DB200_OK;
}

int db200_sketcher_add_record(db200_sketcher *h, uint32_t slot, const char *bases, uint64_t len) {
    if (!h || slot >= h->slots.size() || (!bases && len)) { set_error("db200_sketcher_add_record: bad arguments"); return DB200_EINVAL; }
    auto &s = h->slots[slot];
    try {
        s.bases.insert(s.bases.end(), bases, bases + len);
        s.offs.push_back(s.bases.size());
    } catch (const std::bad_alloc &) { set_error("out of host memory staging a record"); return DB200_ENOMEM; }
    return DB200_OK;
}

int db200_sketcher_finish(db200_sketcher *h, uint32_t slot, uint8_t *registers_out) {
    if (!h || slot >= h->slots.size() || !registers_out) { set_error("db200_sketcher_finish: bad arguments"); return DB200_EINVAL; }
    auto &s = h->slots[slot];
    const uint64_t grb[2] = {0, s.offs.size() - 1};
    // all devices: slots are spread round-robin (one genome per call, so the call itself uses one device)
    const int dev = h->device == DB200_ALL_DEVICES ? (int)(slot % (uint32_t)std::max(logical_device_count(), 1)) : h->device;
    const int rc = db200_sketch_batch(dev, h->p, h->k, h->canon, s.bases.data(), s.offs.data(), s.offs.size() - 1, grb, 1, registers_out);
    s.bases.clear();
    s.offs.assign(1, 0);
    return rc;
}

int db200_sketcher_destroy(db200_sketcher *h) { delete h; return DB200_OK; }

// ---- cardinalities --------------------------------------------------------------------------
int db200_cardinalities(int device, const uint8_t *regs, uint64_t n, int p, int estim, double *out) {
    if ((!regs || !out) && n) { set_error("db200_cardinalities: null argument"); return DB200_EINVAL; }
    if (p < 7 || p > 24) { set_error("cardinalities: p=%d outside the GPU path's range [7,24]", p); return DB200_EUNSUPPORTED; }
    if (estim < 0 || estim > 2) { set_error("unknown estimation method %d", estim); return DB200_EINVAL; }
    if (device == DB200_ALL_DEVICES) {
        const int nd = (int)std::min<uint64_t>((uint64_t)std::max(logical_device_count(), 1), std::max<uint64_t>(n >> 8, 1));
        if (nd > 1) {
            return for_each_device(nd, [&](int d) {
                const uint64_t a = n * (uint64_t)d / (uint64_t)nd, b = n * (uint64_t)(d + 1) / (uint64_t)nd;
                return db200_cardinalities(d, regs + (a << p), b - a, p, estim, out + a);
            });
        }
        device = 0;
    }
    DB200_TRY(check_device(device));
    if (n == 0) return DB200_OK;
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    DB200_TRY(hc.regs.reserve(n << p));
    DB200_TRY(hc.cards.reserve(n * 8));
    DB200_TRY(hc.up.stager.upload(hc.regs.ptr, reinterpret_cast<const char *>(regs), n << p, hc.stream));
    cardinality_kernel<<<(unsigned)n, 128, 0, hc.stream>>>(hc.regs.as<uint32_t>(), p, estim, hc.cards.as<double>());
    DB200_LAUNCHED();
    DB200_CUDA(cudaGetLastError());
    DB200_CUDA(cudaMemcpyAsync(out, hc.cards.ptr, n * 8, cudaMemcpyDeviceToHost, hc.stream));
    DB200_CUDA(cudaStreamSynchronize(hc.stream));
    return DB200_OK;
}

// ---- set operations (union / fold) ----------------------------------------------------------
int db200_union(int device, const uint8_t *regs, uint64_t n, int p, uint8_t *out) {
    if (!out || (!regs && n)) { set_error("db200_union: null argument"); return DB200_EINVAL; }
    if (p < 7 || p > 24) { set_error("union: p=%d outside the GPU path's range [7,24]", p); return DB200_EUNSUPPORTED; }
    if (device == DB200_ALL_DEVICES) device = 0;      // one pass over n * 2^p bytes: PCIe-bound from host memory, one device is enough
    DB200_TRY(check_device(device));
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    const uint64_t m = 1ull << p;
    DB200_TRY(hc.out.reserve(m));
    DB200_CUDA(cudaMemsetAsync(hc.out.ptr, 0, m, hc.stream));
    // sketches stream through the register buffer in slabs (a union of 10^6 sketches need not fit in HBM at once)
    const uint64_t slab = std::max<uint64_t>(1, std::min<uint64_t>(n, (uint64_t(1) << 30) >> p));
    if (n) DB200_TRY(hc.regs.reserve(slab << p));
    const uint32_t m16 = (uint32_t)(m / 16), gx = (m16 + 255) / 256;
    for (uint64_t s0 = 0; s0 < n; s0 += slab) {
        const uint64_t cnt = std::min(slab, n - s0);
        DB200_TRY(hc.up.stager.upload(hc.regs.ptr, reinterpret_cast<const char *>(regs) + (s0 << p), cnt << p, hc.stream));
        const uint64_t want_y = std::max<uint64_t>(1, (uint64_t)g_num_sms(device) * 8 / gx);
        const uint64_t rows_per = std::max<uint64_t>(4, (cnt + want_y - 1) / want_y);
        const unsigned gy = (unsigned)((cnt + rows_per - 1) / rows_per);
        union_kernel<<<dim3(gx, gy), 256, 0, hc.stream>>>(hc.regs.as<uint4>(), cnt, m16, rows_per, hc.out.as<uint32_t>(),
                                                          (gy > 1 || n > slab) ? 1 : 0);
        DB200_LAUNCHED();
    }
    DB200_CUDA(cudaGetLastError());
    DB200_CUDA(cudaMemcpyAsync(out, hc.out.ptr, m, cudaMemcpyDeviceToHost, hc.stream));
    DB200_CUDA(cudaStreamSynchronize(hc.stream));
    return DB200_OK;
}

int db200_compress(int device, const uint8_t *regs, uint64_t n, int p, int new_p, uint8_t *out) {
    if ((!regs || !out) && n) { set_error("db200_compress: null argument"); return DB200_EINVAL; }
    if (p < 7 || p > 24) { set_error("compress: p=%d outside the GPU path's range [7,24]", p); return DB200_EUNSUPPORTED; }
    // hll.h:907-909: equal sizes copy, a larger target throws
    if (new_p > p) { set_error("Can't compress to a larger size. Current: %d. Requested new size: %d", p, new_p); return DB200_EINVAL; }
    if (new_p < 1) { set_error("compress: new p=%d must be positive", new_p); return DB200_EINVAL; }
    if (device == DB200_ALL_DEVICES) device = 0;
    DB200_TRY(check_device(device));
    if (n == 0) return DB200_OK;
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    DB200_TRY(hc.regs.reserve(n << p));
    DB200_TRY(hc.out.reserve(n << new_p));
    DB200_TRY(hc.up.stager.upload(hc.regs.ptr, reinterpret_cast<const char *>(regs), n << p, hc.stream));
    const uint64_t total = n << new_p;
    compress_kernel<<<(unsigned)((total + 255) / 256), 256, 0, hc.stream>>>(hc.regs.as<uint8_t>(), n, p, new_p, hc.out.as<uint8_t>());
    DB200_LAUNCHED();
    DB200_CUDA(cudaGetLastError());
    DB200_CUDA(cudaMemcpyAsync(out, hc.out.ptr, total, cudaMemcpyDeviceToHost, hc.stream));
    DB200_CUDA(cudaStreamSynchronize(hc.stream));
    return DB200_OK;
}

// ---- dist plan ------------------------------------------------------------------------------
int db200_dist_plan_create(int device, db200_dist_plan **out) {
    if (!out) { set_error("db200_dist_plan_create: null out"); return DB200_EINVAL; }
    DB200_TRY(check_device(device));
    db200_dist_plan *pl = new db200_dist_plan;
    pl->device = device;
    *out = pl;
    return DB200_OK;
}
int db200_dist_plan_destroy(db200_dist_plan *pl) {
    if (pl) { cudaSetDevice(phys_of(pl->device)); delete pl; }
    return DB200_OK;
}
int db200_dist_plan_prepare_dev(db200_dist_plan *pl, const uint8_t *d_regs, uint64_t n, int p, int estim, void *stream) {
    if (!pl || !d_regs) { set_error("db200_dist_plan_prepare_dev: null argument"); return DB200_EINVAL; }
    DB200_TRY(check_device(pl->device));
    std::lock_guard<std::mutex> lk(pl->mu);
    return plan_prepare(pl, d_regs, n, n, 0, 0, p, estim, (cudaStream_t)stream);
}
int db200_dist_plan_begin_dev(db200_dist_plan *pl, uint64_t n, int p, int estim, int reg_min, int reg_max, void *stream) {
    if (!pl) { set_error("db200_dist_plan_begin_dev: null plan"); return DB200_EINVAL; }
    DB200_TRY(check_device(pl->device));
    std::lock_guard<std::mutex> lk(pl->mu);
    return plan_begin(pl, n, n, 0, 0, p, estim, reg_min, reg_max, (cudaStream_t)stream);
}
int db200_dist_plan_add_rows_dev(db200_dist_plan *pl, const uint8_t *d_regs, uint64_t row_begin, uint64_t nrows, void *stream) {
    if (!pl || !d_regs) { set_error("db200_dist_plan_add_rows_dev: null argument"); return DB200_EINVAL; }
    DB200_TRY(check_device(pl->device));
    std::lock_guard<std::mutex> lk(pl->mu);
    return plan_add_rows(pl, d_regs, row_begin, nrows, (cudaStream_t)stream);
}
int db200_dist_plan_finish_dev(db200_dist_plan *pl) {
    if (!pl) { set_error("db200_dist_plan_finish_dev: null plan"); return DB200_EINVAL; }
    DB200_TRY(check_device(pl->device));
    std::lock_guard<std::mutex> lk(pl->mu);
    return plan_finish(pl);
}
int db200_dist_plan_run_symmetric_dev(db200_dist_plan *pl, const db200_dist_params *prm, uint64_t row_begin, uint64_t row_end, float *d_out,
                                      void *stream) {
    if (!pl || !prm || !d_out) { set_error("db200_dist_plan_run_symmetric_dev: null argument"); return DB200_EINVAL; }
    DB200_TRY(check_device(pl->device));
    std::lock_guard<std::mutex> lk(pl->mu);
    if (row_end > pl->nrows) row_end = pl->nrows;
    if (row_begin > row_end) { set_error("row_begin > row_end"); return DB200_EINVAL; }
    return plan_run(pl, prm, 0, row_begin, row_end, 0, 0, d_out, (cudaStream_t)stream);
}
int db200_dist_plan_run_rect_dev(db200_dist_plan *pl, const db200_dist_params *prm, uint64_t nr, uint64_t nq, float *d_out, void *stream) {
    if (!pl || !prm || !d_out) { set_error("db200_dist_plan_run_rect_dev: null argument"); return DB200_EINVAL; }
    DB200_TRY(check_device(pl->device));
    std::lock_guard<std::mutex> lk(pl->mu);
    if (pl->qbase == 0) {
        // plan prepared through prepare_dev on a plain [refs; queries] matrix: queries must start on a panel boundary
        if (nr % DT != 0 || nr + nq != pl->nrows) { set_error("rect run on a symmetric plan needs nr %% %d == 0 and nr + nq == n", DT); return DB200_EINVAL; }
        pl->qbase = nr;
        const int rc = plan_run(pl, prm, 1, 0, 0, nr, nq, d_out, (cudaStream_t)stream);
        pl->qbase = 0;
        return rc;
    }
    return plan_run(pl, prm, 1, 0, 0, nr, nq, d_out, (cudaStream_t)stream);
}
int db200_dist_plan_run_knn_dev(db200_dist_plan *pl, const db200_dist_params *prm, uint64_t nr, uint64_t nq, uint32_t nneighbors,
                                db200_neighbor *d_out, void *stream) {
    if (!pl || !prm || !d_out) { set_error("db200_dist_plan_run_knn_dev: null argument"); return DB200_EINVAL; }
    DB200_TRY(check_device(pl->device));
    std::lock_guard<std::mutex> lk(pl->mu);
    static_assert(sizeof(Neighbor) == sizeof(db200_neighbor), "neighbour layout");
    if (nq == 0) return plan_knn(pl, prm, 0, 0, nneighbors, reinterpret_cast<Neighbor *>(d_out), (cudaStream_t)stream);
    if (pl->qbase == 0) {
        if (nr % DT != 0 || nr + nq != pl->nrows) { set_error("rect run on a symmetric plan needs nr %% %d == 0 and nr + nq == n", DT); return DB200_EINVAL; }
        pl->qbase = nr;
        const int rc = plan_knn(pl, prm, nr, nq, nneighbors, reinterpret_cast<Neighbor *>(d_out), (cudaStream_t)stream);
        pl->qbase = 0;
        return rc;
    }
    return plan_knn(pl, prm, nr, nq, nneighbors, reinterpret_cast<Neighbor *>(d_out), (cudaStream_t)stream);
}
static int knn_symmetric_impl(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm, uint32_t nneighbors,
                              db200_neighbor *out, const double *card) {
    if (!prm || ((!regs || !out) && n)) { set_error("db200_dist_knn_symmetric: null argument"); return DB200_EINVAL; }
    // The retained sets depend on the ascending visiting order (ties at the cut), so per-device partial tables cannot be
    // merged exactly: with DB200_ALL_DEVICES the symmetric table is computed on device 0 (the -Q/-F form shards by query).
    if (device == DB200_ALL_DEVICES) device = 0;
    DB200_TRY(check_device(device));
    if (n == 0) return DB200_OK;
    if (prm->p < 7 || prm->p > 20) { set_error("dist: p=%d outside the GPU path's range [7,20]", prm->p); return DB200_EUNSUPPORTED; }
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    const uint64_t m = 1ull << prm->p;
    DB200_TRY(hc.regs.reserve(n * m));
    DB200_TRY(hc.nbrs.reserve(n * nneighbors * sizeof(db200_neighbor)));
    DB200_TRY(hc.up.stager.upload(hc.regs.ptr, reinterpret_cast<const char *>(regs), n * m, hc.stream));
    DB200_TRY(plan_prepare(hc.plan.get(), hc.regs.as<uint8_t>(), n, n, 0, 0, prm->p, prm->estim, hc.stream));
    DB200_TRY(plan_override_cards(hc.plan.get(), card, n, nullptr, 0, hc.stream));
    DB200_TRY(plan_knn(hc.plan.get(), prm, 0, 0, nneighbors, hc.nbrs.as<Neighbor>(), hc.stream));
    DB200_CUDA(cudaMemcpyAsync(out, hc.nbrs.ptr, n * nneighbors * sizeof(db200_neighbor), cudaMemcpyDeviceToHost, hc.stream));
    DB200_CUDA(cudaStreamSynchronize(hc.stream));
    return DB200_OK;
}
static int knn_rect_impl(int device, const uint8_t *ref_regs, uint64_t nr, const uint8_t *qry_regs, uint64_t nq, const db200_dist_params *prm,
                         uint32_t nneighbors, db200_neighbor *out, const double *card_ref, const double *card_qry) {
    if (!prm || ((!ref_regs || !qry_regs || !out) && nr && nq)) { set_error("db200_dist_knn_rect: null argument"); return DB200_EINVAL; }
    if (device == DB200_ALL_DEVICES) {
        const int nd = (int)std::min<uint64_t>((uint64_t)std::max(logical_device_count(), 1), std::max<uint64_t>(nq / DT, 1));
        if (nd > 1 && nr && prm->p >= 7 && prm->p <= 20) {
            const uint64_t m = 1ull << prm->p;
            return for_each_device(nd, [&](int d) {
                const uint64_t a = nq * (uint64_t)d / (uint64_t)nd, b = nq * (uint64_t)(d + 1) / (uint64_t)nd;
                return knn_rect_impl(d, ref_regs, nr, qry_regs + a * m, b - a, prm, nneighbors, out + a * nneighbors, card_ref,
                                     card_qry ? card_qry + a : nullptr);
            });
        }
        device = 0;
    }
    DB200_TRY(check_device(device));
    if (nr == 0 || nq == 0) return DB200_OK;
    if (prm->p < 7 || prm->p > 20) { set_error("dist: p=%d outside the GPU path's range [7,20]", prm->p); return DB200_EUNSUPPORTED; }
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    const uint64_t m = 1ull << prm->p;
    const uint64_t qbase = (nr + DT - 1) / DT * DT, nrows = qbase + nq;
    DB200_TRY(hc.regs.reserve(nrows * m));
    DB200_TRY(hc.nbrs.reserve(nq * nneighbors * sizeof(db200_neighbor)));
    DB200_TRY(hc.up.stager.upload(hc.regs.ptr, reinterpret_cast<const char *>(ref_regs), nr * m, hc.stream));
    DB200_TRY(hc.up.stager.upload(hc.regs.as<uint8_t>() + qbase * m, reinterpret_cast<const char *>(qry_regs), nq * m, hc.stream));
    DB200_TRY(plan_prepare(hc.plan.get(), hc.regs.as<uint8_t>(), nrows, nr, qbase, nq, prm->p, prm->estim, hc.stream));
    DB200_TRY(plan_override_cards(hc.plan.get(), card_ref, nr, card_qry, nq, hc.stream));
    DB200_TRY(plan_knn(hc.plan.get(), prm, nr, nq, nneighbors, hc.nbrs.as<Neighbor>(), hc.stream));
    DB200_CUDA(cudaMemcpyAsync(out, hc.nbrs.ptr, nq * nneighbors * sizeof(db200_neighbor), cudaMemcpyDeviceToHost, hc.stream));
    DB200_CUDA(cudaStreamSynchronize(hc.stream));
    return DB200_OK;
}
int db200_dist_plan_cardinalities_dev(db200_dist_plan *pl, const double **d_card) {
    if (!pl || !d_card || !pl->ready) { set_error("plan not prepared"); return DB200_EINVAL; }
    *d_card = pl->card.as<double>();
    return DB200_OK;
}
int db200_dist_plan_last_run_info(const db200_dist_plan *pl, uint64_t *pairs, uint64_t *tiles, int *thresholds) {
    if (!pl) { set_error("null plan"); return DB200_EINVAL; }
    if (pairs) *pairs = pl->last_pairs;
    if (tiles) *tiles = pl->last_tiles;
    if (thresholds) *thresholds = pl->K;
    return DB200_OK;
}

// ---- host-pointer all-pairs -----------------------------------------------------------------
static int symmetric_rows_impl(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm, uint64_t row_begin, uint64_t row_end,
                               float *out, const double *card) {
    if (!prm || (!regs && n)) { set_error("db200_dist_symmetric_rows: null argument"); return DB200_EINVAL; }
    if (device == DB200_ALL_DEVICES) {
        // block rows balanced by pair count (row i holds n-1-i pairs); rows are contiguous in distmat order, so every
        // device writes its own slice of `out`
        const uint64_t r0 = std::min(row_begin, n), r1 = std::min(row_end, n);
        auto tri = [n](uint64_t r) { return (r * (2 * n - r - 1)) / 2; };
        const uint64_t npairs = r1 > r0 ? tri(r1) - tri(r0) : 0;
        const int nd = (int)std::min<uint64_t>((uint64_t)std::max(logical_device_count(), 1), std::max<uint64_t>(npairs >> 16, 1));
        if (nd > 1 && row_begin <= row_end) {
            const std::vector<uint64_t> cut = split_by_weight(r1 - r0, nd, [&](uint64_t i) { return tri(r0 + i) - tri(r0); });
            return for_each_device(nd, [&](int d) {
                const uint64_t a = r0 + cut[d], b = r0 + cut[d + 1];
                return symmetric_rows_impl(d, regs, n, prm, a, b, out ? out + (tri(a) - tri(r0)) : nullptr, card);
            });
        }
        device = 0;
    }
    DB200_TRY(check_device(device));
    if (row_end > n) row_end = n;
    if (row_begin > row_end) { set_error("row_begin > row_end"); return DB200_EINVAL; }
    if (n < 2 || row_begin == row_end) return DB200_OK;
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    const uint64_t m = 1ull << prm->p;
    auto tri = [n](uint64_t r) { return (r * (2 * n - r - 1)) / 2; };
    const uint64_t npairs = tri(row_end) - tri(row_begin);
    if (npairs && !out) { set_error("null output"); return DB200_EINVAL; }
    if (prm->p < 7 || prm->p > 20) { set_error("dist: p=%d outside the GPU path's range [7,20]", prm->p); return DB200_EUNSUPPORTED; }
    DB200_TRY(hc.regs.reserve(n * m));
    DB200_TRY(hc.out.reserve(std::max<uint64_t>(npairs, 1) * 4));
    DB200_TRY(hc.up.stager.upload(hc.regs.ptr, reinterpret_cast<const char *>(regs), n * m, hc.stream));
    DB200_TRY(plan_prepare(hc.plan.get(), hc.regs.as<uint8_t>(), n, n, 0, 0, prm->p, prm->estim, hc.stream));
    DB200_TRY(plan_override_cards(hc.plan.get(), card, n, nullptr, 0, hc.stream));
    // Row blocks (equal pair counts, up to NSLOT of them): block b's device->host copy runs on the copy stream while block
    // b+1 computes, so only the last block's transfer is exposed.
    const int nblk = npairs >= (uint64_t(4) << 20) ? db200_dist_plan::NSLOT : 1;
    uint64_t rb = row_begin, boff[db200_dist_plan::NSLOT] = {0}, bcnt[db200_dist_plan::NSLOT] = {0};
    for (int b = 0; b < nblk; ++b) {
        uint64_t re = row_end;
        if (b + 1 < nblk) {
            const uint64_t target = tri(row_begin) + npairs * (uint64_t)(b + 1) / (uint64_t)nblk;
            re = rb;
            while (re < row_end && tri(re) < target) ++re;       // block boundaries are whole rows
        }
        if (re == rb) continue;
        boff[b] = tri(rb) - tri(row_begin); bcnt[b] = tri(re) - tri(rb);
        DB200_TRY(plan_run(hc.plan.get(), prm, 0, rb, re, 0, 0, hc.out.as<float>() + boff[b], hc.stream, b));
        DB200_CUDA(cudaEventRecord(hc.blk_done[b], hc.stream));
        rb = re;
    }
    // every block's kernel is queued by now, so a copy that blocks this thread (pageable `out`) still overlaps the kernels behind it
    for (int b = 0; b < nblk; ++b) {
        if (!bcnt[b]) continue;
        DB200_CUDA(cudaStreamWaitEvent(hc.cstream, hc.blk_done[b], 0));
        DB200_TRY(hc.up.stager.download(reinterpret_cast<char *>(out + boff[b]), hc.out.as<float>() + boff[b], bcnt[b] * 4, hc.cstream));
    }
    DB200_CUDA(cudaStreamSynchronize(hc.stream));
    DB200_CUDA(cudaStreamSynchronize(hc.cstream));
    return DB200_OK;
}

static int rect_impl(int device, const uint8_t *ref_regs, uint64_t nr, const uint8_t *qry_regs, uint64_t nq, const db200_dist_params *prm,
                     float *out, const double *card_ref, const double *card_qry) {
    if (!prm || ((!ref_regs || !qry_regs || !out) && nr && nq)) { set_error("db200_dist_rect: null argument"); return DB200_EINVAL; }
    if (device == DB200_ALL_DEVICES) {
        const int nd = (int)std::min<uint64_t>((uint64_t)std::max(logical_device_count(), 1), std::max<uint64_t>(nq / DT, 1));
        if (nd > 1 && nr && prm->p >= 7 && prm->p <= 20) {
            const uint64_t m = 1ull << prm->p;
            return for_each_device(nd, [&](int d) {
                const uint64_t a = nq * (uint64_t)d / (uint64_t)nd, b = nq * (uint64_t)(d + 1) / (uint64_t)nd;
                return rect_impl(d, ref_regs, nr, qry_regs + a * m, b - a, prm, out + a * nr, card_ref, card_qry ? card_qry + a : nullptr);
            });
        }
        device = 0;
    }
    DB200_TRY(check_device(device));
    if (nr == 0 || nq == 0) return DB200_OK;
    if (prm->p < 7 || prm->p > 20) { set_error("dist: p=%d outside the GPU path's range [7,20]", prm->p); return DB200_EUNSUPPORTED; }
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    const uint64_t m = 1ull << prm->p;
    const uint64_t qbase = (nr + DT - 1) / DT * DT, nrows = qbase + nq;
    DB200_TRY(hc.regs.reserve(nrows * m));
    DB200_TRY(hc.out.reserve(nr * nq * 4));
    DB200_TRY(hc.up.stager.upload(hc.regs.ptr, reinterpret_cast<const char *>(ref_regs), nr * m, hc.stream));
    DB200_TRY(hc.up.stager.upload(hc.regs.as<uint8_t>() + qbase * m, reinterpret_cast<const char *>(qry_regs), nq * m, hc.stream));
    DB200_TRY(plan_prepare(hc.plan.get(), hc.regs.as<uint8_t>(), nrows, nr, qbase, nq, prm->p, prm->estim, hc.stream));
    DB200_TRY(plan_override_cards(hc.plan.get(), card_ref, nr, card_qry, nq, hc.stream));
    DB200_TRY(plan_run(hc.plan.get(), prm, 1, 0, 0, nr, nq, hc.out.as<float>(), hc.stream));
    DB200_TRY(hc.up.stager.download(reinterpret_cast<char *>(out), hc.out.ptr, nr * nq * 4, hc.stream));
    DB200_CUDA(cudaStreamSynchronize(hc.stream));
    return DB200_OK;
}

} // extern "C"

// Row-block streaming form of the symmetric all-pairs call: the packed triangle never exists in full on either side.  Blocks of
// whole rows (about `block_pairs` values each) go kernel -> device buffer -> page-locked host buffer -> callback, three blocks
// in flight, so the host's writer runs while the next blocks are being computed and copied.
static int symmetric_stream_impl(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm, uint64_t row_begin, uint64_t row_end,
                                 uint64_t block_pairs, db200_rows_cb cb, void *ud, const double *card) {
    if (!prm || !cb || (!regs && n)) { set_error("db200_dist_symmetric_stream: null argument"); return DB200_EINVAL; }
    if (row_end > n) row_end = n;
    if (row_begin > row_end) { set_error("row_begin > row_end"); return DB200_EINVAL; }
    auto tri = [n](uint64_t r) { return (r * (2 * n - r - 1)) / 2; };
    if (device == DB200_ALL_DEVICES) {
        const uint64_t npairs = tri(row_end) - tri(row_begin);
        const int nd = (int)std::min<uint64_t>((uint64_t)std::max(logical_device_count(), 1), std::max<uint64_t>(npairs >> 16, 1));
        if (nd > 1) {
            const uint64_t r0 = row_begin;
            const std::vector<uint64_t> cut = split_by_weight(row_end - r0, nd, [&](uint64_t i) { return tri(r0 + i) - tri(r0); });
            return for_each_device(nd, [&](int d) {
                return symmetric_stream_impl(d, regs, n, prm, r0 + cut[d], r0 + cut[d + 1], block_pairs, cb, ud, card);
            });
        }
        device = 0;
    }
    DB200_TRY(check_device(device));
    if (n < 2 || row_begin == row_end) return DB200_OK;
    if (prm->p < 7 || prm->p > 20) { set_error("dist: p=%d outside the GPU path's range [7,20]", prm->p); return DB200_EUNSUPPORTED; }
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    const uint64_t m = 1ull << prm->p;
    if (block_pairs == 0) block_pairs = uint64_t(8) << 20;
    block_pairs = std::max<uint64_t>(block_pairs, n);                  // at least one whole row
    constexpr int NS = HostCtx::NSTREAM;
    // a block may overshoot the target by one row
    const uint64_t slot_vals = block_pairs + n;
    if (hc.stream_cap < slot_vals * 4) {
        for (int i = 0; i < NS; ++i) {
            if (hc.stream_host[i]) { cudaFreeHost(hc.stream_host[i]); hc.stream_host[i] = nullptr; }
            DB200_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&hc.stream_host[i]), slot_vals * 4, cudaHostAllocDefault));
            if (!hc.stream_copied[i]) DB200_CUDA(cudaEventCreateWithFlags(&hc.stream_copied[i], cudaEventDisableTiming));
        }
        hc.stream_cap = slot_vals * 4;
    }
    DB200_TRY(hc.regs.reserve(n * m));
    DB200_TRY(hc.out.reserve(slot_vals * 4 * NS));
    DB200_TRY(hc.up.stager.upload(hc.regs.ptr, reinterpret_cast<const char *>(regs), n * m, hc.stream));
    DB200_TRY(plan_prepare(hc.plan.get(), hc.regs.as<uint8_t>(), n, n, 0, 0, prm->p, prm->estim, hc.stream));
    DB200_TRY(plan_override_cards(hc.plan.get(), card, n, nullptr, 0, hc.stream));
    struct Blk { uint64_t rb, re, cnt; };
    Blk ring[NS];
    uint64_t issued = 0, delivered = 0;
    auto deliver = [&]() -> int {
        const Blk &b = ring[delivered % NS];
        DB200_CUDA(cudaEventSynchronize(hc.stream_copied[delivered % NS]));
        const int rc = cb(ud, b.rb, b.re, reinterpret_cast<const float *>(hc.stream_host[delivered % NS]), b.cnt);
        ++delivered;
        if (rc != 0) { set_error("db200_dist_symmetric_stream: the callback returned %d for rows [%llu,%llu)", rc, (unsigned long long)b.rb, (unsigned long long)b.re); return DB200_EINVAL; }
        return DB200_OK;
    };
    for (uint64_t rb = row_begin; rb < row_end;) {
        uint64_t re = rb + 1;
        while (re < row_end && tri(re) - tri(rb) < block_pairs) ++re;
        const uint64_t cnt = tri(re) - tri(rb);
        if (issued - delivered == NS) DB200_TRY(deliver());            // the slot about to be reused must have been handed over
        const int sl = (int)(issued % NS);
        float *d_out = hc.out.as<float>() + (uint64_t)sl * slot_vals;
        DB200_TRY(plan_run(hc.plan.get(), prm, 0, rb, re, 0, 0, d_out, hc.stream, (int)(issued % db200_dist_plan::NSLOT)));
        DB200_CUDA(cudaEventRecord(hc.blk_done[issued % db200_dist_plan::NSLOT], hc.stream));
        DB200_CUDA(cudaStreamWaitEvent(hc.cstream, hc.blk_done[issued % db200_dist_plan::NSLOT], 0));
        if (cnt) DB200_CUDA(cudaMemcpyAsync(hc.stream_host[sl], d_out, cnt * 4, cudaMemcpyDeviceToHost, hc.cstream));
        DB200_CUDA(cudaEventRecord(hc.stream_copied[sl], hc.cstream));
        ring[sl] = Blk{rb, re, cnt};
        ++issued;
        rb = re;
    }
    while (delivered < issued) DB200_TRY(deliver());
    DB200_CUDA(cudaStreamSynchronize(hc.stream));
    DB200_CUDA(cudaStreamSynchronize(hc.cstream));
    return DB200_OK;
}

// ---- public all-pairs entry points (cached cardinalities travel in the params: prm->card / prm->card_queries) ----------
extern "C" {

int db200_dist_symmetric_rows(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm, uint64_t row_begin, uint64_t row_end,
                              float *out) {
    return symmetric_rows_impl(device, regs, n, prm, row_begin, row_end, out, prm ? prm->card : nullptr);
}
int db200_dist_symmetric(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm, float *out) {
    return db200_dist_symmetric_rows(device, regs, n, prm, 0, n, out);
}
int db200_dist_symmetric_stream(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm, uint64_t row_begin, uint64_t row_end,
                                uint64_t block_pairs, db200_rows_cb cb, void *user) {
    return symmetric_stream_impl(device, regs, n, prm, row_begin, row_end, block_pairs, cb, user, prm ? prm->card : nullptr);
}
int db200_dist_rect(int device, const uint8_t *ref_regs, uint64_t nr, const uint8_t *qry_regs, uint64_t nq, const db200_dist_params *prm,
                    float *out) {
    return rect_impl(device, ref_regs, nr, qry_regs, nq, prm, out, prm ? prm->card : nullptr, prm ? prm->card_queries : nullptr);
}
int db200_dist_plan_run_knn_rows_dev(db200_dist_plan *pl, const db200_dist_params *prm, uint64_t row_begin, uint64_t row_end, uint32_t nneighbors,
                                     db200_neighbor *d_out, void *stream) {
    if (!pl || !prm || !d_out) { set_error("db200_dist_plan_run_knn_rows_dev: null argument"); return DB200_EINVAL; }
    if (row_begin > row_end) { set_error("row_begin > row_end"); return DB200_EINVAL; }
    DB200_TRY(check_device(pl->device));
    std::lock_guard<std::mutex> lk(pl->mu);
    return plan_knn(pl, prm, 0, 0, nneighbors, reinterpret_cast<Neighbor *>(d_out), (cudaStream_t)stream, row_begin, row_end);
}
int db200_dist_knn_symmetric(int device, const uint8_t *regs, uint64_t n, const db200_dist_params *prm, uint32_t nneighbors,
                             db200_neighbor *out) {
    return knn_symmetric_impl(device, regs, n, prm, nneighbors, out, prm ? prm->card : nullptr);
}
int db200_dist_knn_rect(int device, const uint8_t *ref_regs, uint64_t nr, const uint8_t *qry_regs, uint64_t nq, const db200_dist_params *prm,
                        uint32_t nneighbors, db200_neighbor *out) {
    return knn_rect_impl(device, ref_regs, nr, qry_regs, nq, prm, nneighbors, out, prm ? prm->card : nullptr, prm ? prm->card_queries : nullptr);
}

} // extern "C"

// ---- device-side FASTA parsing -----------------------------------------------------------------------------------
// Raw file text goes up in 64 MiB chunks; each chunk is summarised, chained and emitted (fasta.cuh) while the next one is in
// flight, and every genome group is sketched as soon as its last chunk has been emitted.  Genome g's bases land in a window
// of the position space sized by the raw length of its files (an upper bound); the work items laid over the window are
// clipped on the device to what the parse produced.
static int sketch_fasta_one(int device, int p, int k, int canon, const char *text, const uint64_t *file_off, const uint64_t *file_len,
                            uint64_t nfiles, const uint64_t *genome_file_begin, uint64_t ngenomes, uint8_t *registers_out,
                            uint8_t *file_status_out) {
    DB200_TRY(check_device(device));
    HostCtx &hc = host_ctx(device);
    std::lock_guard<std::mutex> lk(hc.mu);
    DB200_TRY(hc.init(device));
    DB200_TRY(hc.up.init());
    Uploader &up = hc.up;
    cudaStream_t stream = hc.stream;
    db200_packed_genomes *pg = hc.store.get();
    const uint64_t B = FA_BLOCK;
    const uint64_t text_end = nfiles ? file_off[nfiles - 1] + file_len[nfiles - 1] : 0;
    const uint64_t nblk_text = (text_end + B - 1) / B;
    // ---- tables: first block / length / window start / genome of every file; window of every genome
    std::vector<uint64_t> fblk(nfiles), gpos0(nfiles, ~0ull), gbase(ngenomes + 1, 0);
    std::vector<uint32_t> fgen(nfiles);
    uint64_t posn = 0;
    for (uint64_t g = 0; g < ngenomes; ++g) {
        gbase[g] = posn;
        uint64_t raw = 0;
        for (uint64_t f = genome_file_begin[g]; f < genome_file_begin[g + 1]; ++f) {
            fblk[f] = file_off[f] / B;
            fgen[f] = (uint32_t)g;
            if (f == genome_file_begin[g]) gpos0[f] = posn;
            raw += file_len[f];
        }
        posn = (posn + raw + 63 + 64) & ~63ull;     // at least one all-invalid block between genomes
    }
    gbase[ngenomes] = posn;
    pg->device = device; pg->k = k; pg->nbases = posn; pg->ngenomes = ngenomes; pg->kmers = 0;
    pg->nblk = posn / 64 + 1;
    DB200_TRY(pg->bases2.reserve(pg->nblk * 16));
    DB200_TRY(pg->nb.reserve(pg->nblk * 8));
    DB200_TRY(pg->st.reserve(pg->nblk * 8));
    DB200_TRY(pg->counter.reserve(64));
    DB200_CUDA(cudaMemsetAsync(pg->bases2.ptr, 0, pg->nblk * 16, stream));
    DB200_CUDA(cudaMemsetAsync(pg->nb.ptr, 0, pg->nblk * 8, stream));
    DB200_CUDA(cudaMemsetAsync(pg->st.ptr, 0, pg->nblk * 8, stream));
    DB200_CUDA(cudaMemsetAsync(pg->counter.ptr, 0, 64, stream));
    const uint64_t bytes = ngenomes << p;
    DB200_TRY(hc.regs.reserve(std::max<uint64_t>(bytes, 16)));
    DB200_CUDA(cudaMemsetAsync(hc.regs.ptr, 0, std::max<uint64_t>(bytes, 16), stream));
    DB200_TRY(hc.fa_text.reserve(std::max<uint64_t>(nblk_text, 1) * B + 64));
    DB200_TRY(hc.fa_sums.reserve(std::max<uint64_t>(nblk_text, 1) * 8));
    DB200_TRY(hc.fa_wsum.reserve(((64ull << 20) / B) * (FA_THREADS / 32) * 8));      // per-warp maps of the chunk in flight
    DB200_TRY(hc.fa_state.reserve(std::max<uint64_t>(nblk_text, 1)));
    DB200_TRY(hc.fa_pos.reserve(std::max<uint64_t>(nblk_text, 1) * 8));
    DB200_TRY(hc.fa_gend.reserve((ngenomes + 1) * 8));
    DB200_TRY(hc.fa_carry.reserve(32));
    DB200_TRY(hc.fa_flags.reserve(std::max<uint64_t>(nfiles, 1) * 4));
    const uint64_t tab_bytes = nfiles * (8 + 8 + 8 + 4);
    DB200_TRY(hc.fa_tab.reserve(std::max<uint64_t>(tab_bytes, 16)));
    uint64_t *d_fblk = hc.fa_tab.as<uint64_t>(), *d_flen = d_fblk + nfiles, *d_gpos0 = d_flen + nfiles;
    uint32_t *d_fgen = reinterpret_cast<uint32_t *>(d_gpos0 + nfiles);
    if (nfiles) {
        DB200_CUDA(cudaMemcpyAsync(d_fblk, fblk.data(), nfiles * 8, cudaMemcpyHostToDevice, stream));
        DB200_CUDA(cudaMemcpyAsync(d_flen, file_len, nfiles * 8, cudaMemcpyHostToDevice, stream));
        DB200_CUDA(cudaMemcpyAsync(d_gpos0, gpos0.data(), nfiles * 8, cudaMemcpyHostToDevice, stream));
        DB200_CUDA(cudaMemcpyAsync(d_fgen, fgen.data(), nfiles * 4, cudaMemcpyHostToDevice, stream));
    }
    // genome_end starts at each window's begin (a genome without sequence bytes clips all its items away)
    DB200_CUDA(cudaMemcpyAsync(hc.fa_gend.ptr, gbase.data(), (ngenomes + 1) * 8, cudaMemcpyHostToDevice, stream));
    DB200_CUDA(cudaMemsetAsync(hc.fa_carry.ptr, 0, 32, stream));
    DB200_CUDA(cudaMemsetAsync(hc.fa_flags.ptr, 0, std::max<uint64_t>(nfiles, 1) * 4, stream));
    // ---- work items over the windows, in up to NGROUP genome groups of similar raw size (as in pack_genomes_impl)
    std::vector<SketchItem> items;
    pg->group_item_begin.assign(1, 0u);
    pg->group_end.clear();                       // here: text offset after which the group is complete
    {
        const uint64_t target_items = (uint64_t)g_num_sms(device) * 16;
        uint64_t chunk = posn / std::max<uint64_t>(target_items, 1);
        chunk = std::min<uint64_t>(std::max<uint64_t>(chunk, 1ull << 16), 1ull << 20);
        uint64_t g0 = 0;
        for (int grp = 0; grp < Uploader::NGROUP && g0 < ngenomes; ++grp) {
            uint64_t g1 = ngenomes;
            if (grp + 1 < Uploader::NGROUP) {
                const uint64_t want = posn * (uint64_t)(grp + 1) / (uint64_t)Uploader::NGROUP;
                g1 = g0 + 1;
                while (g1 < ngenomes && gbase[g1] < want) ++g1;
            }
            std::vector<std::pair<uint32_t, SketchItem>> tmp;
            for (uint64_t g = g0; g < g1; ++g) {
                const uint64_t gs = gbase[g], ge = gbase[g + 1] - 64;     // window without the separating block
                if (ge <= gs) continue;
                const uint64_t nch = (ge - gs + chunk - 1) / chunk;
                const uint64_t step = (((ge - gs + nch - 1) / nch) + 63) & ~63ull;
                uint32_t ci = 0;
                for (uint64_t st = gs; st < ge; st += step, ++ci) tmp.push_back({ci, SketchItem{st, std::min(ge, st + step), (uint32_t)g, 0}});
            }
            std::stable_sort(tmp.begin(), tmp.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
            for (auto &t : tmp) items.push_back(t.second);
            pg->group_item_begin.push_back((uint32_t)items.size());
            // the group is complete once the last byte of its last file is on the device and emitted
            const uint64_t lf = genome_file_begin[g1] - 1;
            pg->group_end.push_back(genome_file_begin[g1] > genome_file_begin[g0] ? file_off[lf] + file_len[lf] : 0);
            g0 = g1;
        }
    }
    pg->nitems = (uint32_t)items.size();
    if (!items.empty()) {
        DB200_TRY(pg->items.reserve(items.size() * sizeof(SketchItem)));
        DB200_CUDA(cudaMemcpyAsync(pg->items.ptr, items.data(), items.size() * sizeof(SketchItem), cudaMemcpyHostToDevice, stream));
    }
    // ---- chunks
    uint64_t CH = 64ull << 20;                   // a multiple of FA_BLOCK
    if (const char *cenv = std::getenv("DB200_FASTA_CHUNK")) {   // testing knob: chunk boundaries every few blocks
        const uint64_t v = std::strtoull(cenv, nullptr, 10) / B * B;
        if (v) CH = std::min(v, CH);
    }
    const uint64_t nchunks = (text_end + CH - 1) / CH;
    DB200_CUDA(cudaEventRecord(up.packed[0], stream));
    DB200_CUDA(cudaStreamWaitEvent(up.cs, up.packed[0], 0));
    uint8_t *d_text = hc.fa_text.as<uint8_t>();
    size_t next_group = 0;
    for (uint64_t c = 0; c < nchunks; ++c) {
        const int b = (int)(c & 1);
        const uint64_t off = c * CH, len = std::min<uint64_t>(CH, text_end - off);
        DB200_TRY(up.stager.upload(d_text + off, text + off, len, up.cs));
        DB200_CUDA(cudaEventRecord(up.copied[b], up.cs));
        DB200_CUDA(cudaStreamWaitEvent(stream, up.copied[b], 0));
        const uint64_t blk0 = off / B, nb = (len + B - 1) / B;
        const uint32_t next_byte = off + len < text_end ? (uint32_t)(uint8_t)text[off + len] : (uint32_t)'\n';
        fa_summary_kernel<<<(unsigned)nb, FA_THREADS, 0, stream>>>(d_text, blk0, d_fblk, d_flen, (uint32_t)nfiles, off + len, next_byte, hc.fa_sums.as<FaSum>(), hc.fa_wsum.as<FaSum>());
        DB200_LAUNCHED();
        fa_chain_kernel<<<1, FA_CHAIN_THREADS, 0, stream>>>(hc.fa_sums.as<FaSum>(), blk0, nb, d_fblk, d_gpos0, d_fgen, (uint32_t)nfiles, hc.fa_carry.as<uint64_t>(),
                                               hc.fa_state.as<uint8_t>(), hc.fa_pos.as<uint64_t>(), hc.fa_gend.as<uint64_t>());
        DB200_LAUNCHED();
        fa_emit_kernel<<<(unsigned)nb, FA_THREADS, 0, stream>>>(d_text, blk0, d_fblk, d_flen, (uint32_t)nfiles, off + len, next_byte, hc.fa_state.as<uint8_t>(),
                                                                hc.fa_pos.as<uint64_t>(), hc.fa_sums.as<FaSum>(), hc.fa_wsum.as<FaSum>(), pg->bases2.as<uint32_t>(), pg->nb.as<uint32_t>(),
                                                                pg->st.as<uint32_t>(), hc.fa_flags.as<uint32_t>());
        DB200_LAUNCHED();
        while (next_group < pg->group_end.size() && pg->group_end[next_group] <= off + len) {
            const uint32_t ib = pg->group_item_begin[next_group], ie = pg->group_item_begin[next_group + 1];
            if (ie > ib) {
                fa_clip_items_kernel<<<(ie - ib + 255) / 256, 256, 0, stream>>>(pg->items.as<SketchItem>() + ib, ie - ib, hc.fa_gend.as<uint64_t>());
                DB200_LAUNCHED();
                DB200_CUDA(cudaEventRecord(up.group_packed[next_group], stream));
                DB200_CUDA(cudaStreamWaitEvent(up.ss, up.group_packed[next_group], 0));
                DB200_TRY(sketch_launch(pg, p, canon, hc.regs.as<uint8_t>(), up.ss, ib, ie - ib, (int)next_group));
            }
            ++next_group;
        }
    }
    // groups made of empty files only (no chunk ever reached them)
    while (next_group < pg->group_end.size()) {
        const uint32_t ib = pg->group_item_begin[next_group], ie = pg->group_item_begin[next_group + 1];
        if (ie > ib) {
            fa_clip_items_kernel<<<(ie - ib + 255) / 256, 256, 0, stream>>>(pg->items.as<SketchItem>() + ib, ie - ib, hc.fa_gend.as<uint64_t>());
            DB200_LAUNCHED();
            DB200_CUDA(cudaEventRecord(up.group_packed[next_group], stream));
            DB200_CUDA(cudaStreamWaitEvent(up.ss, up.group_packed[next_group], 0));
            DB200_TRY(sketch_launch(pg, p, canon, hc.regs.as<uint8_t>(), up.ss, ib, ie - ib, (int)next_group));
        }
        ++next_group;
    }
    DB200_CUDA(cudaGetLastError());
    {
        const cudaError_t es = cudaStreamSynchronize(up.ss);
        if (es != cudaSuccess) { set_error("sketch (device-parsed FASTA): %s", cudaGetErrorString(es)); return DB200_ECUDA; }
    }
    std::vector<uint32_t> flags(nfiles, 0);
    if (nfiles) DB200_CUDA(cudaMemcpyAsync(flags.data(), hc.fa_flags.ptr, nfiles * 4, cudaMemcpyDeviceToHost, stream));
    if (bytes) DB200_CUDA(cudaMemcpyAsync(registers_out, hc.regs.ptr, bytes, cudaMemcpyDeviceToHost, stream));
    const cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("sketch (device-parsed FASTA): %s", cudaGetErrorString(e)); return DB200_ECUDA; }
    if (file_status_out) for (uint64_t f = 0; f < nfiles; ++f) file_status_out[f] = (uint8_t)(flags[f] & 1u);
    return DB200_OK;
}

extern "C" int db200_sketch_fasta_batch(int device, int p, int k, int canon, const char *text, const uint64_t *file_off, const uint64_t *file_len,
                                        uint64_t nfiles, const uint64_t *genome_file_begin, uint64_t ngenomes, uint8_t *registers_out,
                                        uint8_t *file_status_out) {
    if (!registers_out && ngenomes) { set_error("db200_sketch_fasta_batch: null output"); return DB200_EINVAL; }
    if (!genome_file_begin || ((!file_off || !file_len || !text) && nfiles)) { set_error("db200_sketch_fasta_batch: null argument"); return DB200_EINVAL; }
    if (k < 1 || k > 32) { set_error("sketch: k=%d outside [1,32]", k); return DB200_EUNSUPPORTED; }
    if (p < 7 || p > 24) { set_error("sketch: p=%d outside the GPU path's range [7,24]", p); return DB200_EUNSUPPORTED; }
    if (ngenomes >= (1ull << 32) || nfiles >= (1ull << 32)) { set_error("too many genomes / files"); return DB200_EINVAL; }
    if (genome_file_begin[0] != 0 || genome_file_begin[ngenomes] != nfiles) { set_error("db200_sketch_fasta_batch: genome_file_begin must run from 0 to nfiles"); return DB200_EINVAL; }
    for (uint64_t g = 0; g < ngenomes; ++g)
        if (genome_file_begin[g] > genome_file_begin[g + 1]) { set_error("db200_sketch_fasta_batch: genome_file_begin must ascend"); return DB200_EINVAL; }
    for (uint64_t f = 0; f < nfiles; ++f) {
        if (file_off[f] % FA_BLOCK) { set_error("db200_sketch_fasta_batch: file %llu does not start on a multiple of %d bytes", (unsigned long long)f, FA_BLOCK); return DB200_EINVAL; }
        if (f + 1 < nfiles && (file_off[f + 1] <= file_off[f] || file_off[f] + file_len[f] > file_off[f + 1])) {
            set_error("db200_sketch_fasta_batch: files must ascend, one block at least apart, without overlap (file %llu)", (unsigned long long)f);
            return DB200_EINVAL;
        }
    }
    if (device == DB200_ALL_DEVICES && logical_device_count() > 1 && ngenomes > 1) {
        const int nd = (int)std::min<uint64_t>((uint64_t)logical_device_count(), ngenomes);
        std::vector<uint64_t> raw(ngenomes + 1, 0);
        for (uint64_t g = 0; g < ngenomes; ++g) {
            raw[g + 1] = raw[g] + 1;
            for (uint64_t f = genome_file_begin[g]; f < genome_file_begin[g + 1]; ++f) raw[g + 1] += file_len[f];
        }
        const std::vector<uint64_t> cut = split_by_weight(ngenomes, nd, [&](uint64_t g) { return raw[g]; });
        return for_each_device(nd, [&](int d) {
            const uint64_t g0 = cut[d], g1 = cut[d + 1];
            if (g1 <= g0) return (int)DB200_OK;
            const uint64_t f0 = genome_file_begin[g0], f1 = genome_file_begin[g1];
            // the device's share as a batch of its own: offsets relative to its first file's block
            const uint64_t base = f1 > f0 ? file_off[f0] : 0;
            std::vector<uint64_t> fo(f1 - f0), gfb(g1 - g0 + 1);
            for (uint64_t f = f0; f < f1; ++f) fo[f - f0] = file_off[f] - base;
            for (uint64_t g = g0; g <= g1; ++g) gfb[g - g0] = genome_file_begin[g] - f0;
            return sketch_fasta_one(d, p, k, canon, text + base, fo.data(), file_len + f0, f1 - f0, gfb.data(), g1 - g0,
                                    registers_out + (g0 << p), file_status_out ? file_status_out + f0 : nullptr);
        });
    }
    return sketch_fasta_one(device == DB200_ALL_DEVICES ? 0 : device, p, k, canon, text, file_off, file_len, nfiles, genome_file_begin, ngenomes,
                            registers_out, file_status_out);
}

// Genomes are independent units: contiguous ranges of genomes, balanced by their bases, one range per device.
static int sketch_batch_all(int p, int k, int canon, const char *bases, const uint64_t *rec_offsets, uint64_t nrecords,
                            const uint64_t *genome_rec_begin, uint64_t ngenomes, uint8_t *registers_out) {
    (void)nrecords;
    const int nd = (int)std::min<uint64_t>((uint64_t)logical_device_count(), ngenomes);
    const uint64_t b0 = rec_offsets[genome_rec_begin[0]];
    const std::vector<uint64_t> cut = split_by_weight(ngenomes, nd, [&](uint64_t g) { return rec_offsets[genome_rec_begin[g]] - b0 + g; });
    return for_each_device(nd, [&](int d) {
        const uint64_t g0 = cut[d], g1 = cut[d + 1];
        if (g1 <= g0) return (int)DB200_OK;
        const uint64_t r0 = genome_rec_begin[g0], r1 = genome_rec_begin[g1];
        std::vector<uint64_t> grb(g1 - g0 + 1);
        for (uint64_t g = g0; g <= g1; ++g) grb[g - g0] = genome_rec_begin[g] - r0;
        return db200_sketch_batch(d, p, k, canon, bases, rec_offsets + r0, r1 - r0, grb.data(), g1 - g0,
                                  registers_out ? registers_out + (g0 << p) : nullptr);
    });
}
