// tb_runtime.cu -- runtime of libtoastb200.so: errors, device selection, the host->device
// pointer table (replaces the reference's OmpManager, accelerator.cpp:228-745) and the
// per-call resolver for the three `mem` modes.
#include "tb_runtime.cuh"

#include <atomic>

namespace tbr {

static thread_local std::string g_err;
static thread_local int g_code = 0;
static std::atomic<int64_t> g_launches{0};
static std::mutex g_mu;
static std::unordered_map<const void *, TableEntry> g_table;
static size_t g_table_bytes = 0;
static int g_device = -1;     // -1: whatever device is current in this thread
static int g_disabled = 0;
static int g_probe = -1;      // -1 unknown, 0 no device, 1 ok
static int g_sm_count = 0;

void set_error(int code, const std::string &msg) {
    g_code = code;
    g_err = msg;
}
int last_code() { return g_code; }
const std::string &last_msg() { return g_err; }

void count_launch(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launches() { return g_launches.load(); }

static int probe() {
    if (g_probe < 0) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0) {
            cudaGetLastError();
            g_probe = 0;
        } else {
            g_probe = 1;
        }
    }
    return g_probe;
}

bool enabled() { return probe() == 1 && !g_disabled; }

void require_device() {
    if (!enabled()) {
        throw Error{TB_ERR_NO_DEVICE,
                    "libtoastb200: no usable CUDA device (there is no CPU fallback on this path)"};
    }
    if (g_device >= 0) {
        int cur = -1;
        TB_CUDA(cudaGetDevice(&cur));
        if (cur != g_device) TB_CUDA(cudaSetDevice(g_device));
    }
    if (g_sm_count == 0) {
        int dev = 0;
        TB_CUDA(cudaGetDevice(&dev));
        TB_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
}

int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

int assign_device(int node_procs, int node_rank, double mem_gb, int disabled) {
    (void)node_procs;
    (void)mem_gb;
    g_disabled = disabled;
    if (disabled) return TB_OK;
    if (probe() != 1) return TB_OK; // like the reference: silently stay on "host" = no accel
    int n = 0;
    TB_CUDA(cudaGetDeviceCount(&n));
    g_device = node_rank % n;
    TB_CUDA(cudaSetDevice(g_device));
    g_sm_count = 0;
    return TB_OK;
}

int get_device() {
    if (!enabled()) return -1;
    if (g_device >= 0) return g_device;
    int cur = -1;
    cudaGetDevice(&cur);
    return cur;
}

void *table_lookup(const void *host) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_table.find(host);
    return it == g_table.end() ? nullptr : it->second.dev;
}

static TableEntry *table_find(const void *host) {
    auto it = g_table.find(host);
    return it == g_table.end() ? nullptr : &it->second;
}

// ---- Resolver --------------------------------------------------------------------------------

Resolver::Resolver(int mem, void *stream) : mem_(mem), stream_((cudaStream_t)stream) {
    require_device();
    TB_REQUIRE(mem == TB_MEM_HOST || mem == TB_MEM_DEVICE || mem == TB_MEM_TABLE,
               "invalid `mem` mode");
    // host-buffer modes behave like the reference: the call returns when the result is there
    if (mem != TB_MEM_DEVICE) need_sync_ = true;
}

Resolver::~Resolver() {
    if (!finished_) {
        // error path: release temporaries without copying back
        for (void *t : temps_) cudaFreeAsync(t, stream_);
        cudaStreamSynchronize(stream_);
    }
}

void *Resolver::scratch(size_t nbytes) {
    void *d = nullptr;
    TB_CUDA(cudaMallocAsync(&d, nbytes ? nbytes : 1, stream_));
    temps_.push_back(d);
    return d;
}

void *Resolver::upload(const void *p, size_t nbytes) {
    if (p == nullptr) return nullptr;
    void *d = scratch(nbytes);
    if (nbytes) TB_CUDA(cudaMemcpyAsync(d, p, nbytes, cudaMemcpyHostToDevice, stream_));
    return d;
}

void *Resolver::resolve(void *p, size_t nbytes, bool copy_in, bool copy_out) {
    if (p == nullptr) return nullptr;
    if (mem_ == TB_MEM_DEVICE) return p;
    if (mem_ == TB_MEM_TABLE) {
        std::lock_guard<std::mutex> lk(g_mu);
        TableEntry *e = table_find(p);
        if (e == nullptr) {
            char buf[128];
            snprintf(buf, sizeof(buf), "host buffer %p is not present on the device", p);
            throw Error{TB_ERR_NOT_PRESENT, buf};
        }
        if (e->nbytes < nbytes) {
            throw Error{TB_ERR_ARG, "device copy of '" + e->name + "' is smaller than the buffer"};
        }
        return e->dev;
    }
    void *d = scratch(nbytes);
    if (copy_in && nbytes) TB_CUDA(cudaMemcpyAsync(d, p, nbytes, cudaMemcpyHostToDevice, stream_));
    if (copy_out) backs_.push_back(Back{d, p, nbytes, false});
    return d;
}

void Resolver::finish() {
    for (const Back &b : backs_) {
        if (b.nbytes)
            TB_CUDA(cudaMemcpyAsync(b.host, b.dev, b.nbytes, cudaMemcpyDeviceToHost, stream_));
    }
    for (void *t : temps_) TB_CUDA(cudaFreeAsync(t, stream_));
    temps_.clear();
    finished_ = true;
    TB_CUDA(cudaGetLastError());
    if (need_sync_) TB_CUDA(cudaStreamSynchronize(stream_));
}

} // namespace tbr

// ---------------------------------------------------------------------------------------------
// C ABI: runtime + memory table
// ---------------------------------------------------------------------------------------------
namespace tbr {
const std::string &last_msg();
int64_t launches();
bool enabled();
int assign_device(int, int, double, int);
int get_device();
} // namespace tbr

extern "C" {

const char *tb_last_error(void) { return tbr::last_msg().c_str(); }
const char *tb_version(void) { return "toast_b200 0.1 (sm_100a)"; }
int tb_accel_enabled(void) { return tbr::enabled() ? 1 : 0; }

int tb_accel_assign_device(int node_procs, int node_rank, double mem_gb, int disabled) {
    TB_API_BEGIN
    tbr::assign_device(node_procs, node_rank, mem_gb, disabled);
    TB_API_END
}

int tb_accel_get_device(void) { return tbr::get_device(); }

int tb_device_synchronize(void) {
    TB_API_BEGIN
    tbr::require_device();
    TB_CUDA(cudaDeviceSynchronize());
    TB_API_END
}

int64_t tb_launch_count(void) { return tbr::launches(); }

int tb_accel_present(const void *host, size_t nbytes) {
    (void)nbytes;
    if (!tbr::enabled()) return 0;
    return tbr::table_lookup(host) != nullptr ? 1 : 0;
}

int tb_accel_create(const void *host, size_t nbytes, const char *name) {
    TB_API_BEGIN
    tbr::require_device();
    std::lock_guard<std::mutex> lk(tbr::g_mu);
    if (tbr::g_table.count(host)) {
        // accelerator.cpp:339-347 -- creating twice is an error
        throw tbr::Error{TB_ERR_ALREADY_PRESENT,
                         std::string("accel_create: buffer '") + (name ? name : "") +
                             "' already present on device"};
    }
    void *d = nullptr;
    TB_CUDA(cudaMalloc(&d, nbytes ? nbytes : 1));
    tbr::g_table[host] = tbr::TableEntry{d, nbytes, name ? name : ""};
    tbr::g_table_bytes += nbytes;
    TB_API_END
}

static tbr::TableEntry *need_entry(const void *host, size_t nbytes, const char *name,
                                   const char *what) {
    tbr::TableEntry *e = tbr::table_find(host);
    if (e == nullptr) {
        throw tbr::Error{TB_ERR_NOT_PRESENT, std::string(what) + ": buffer '" +
                                                 (name ? name : "") + "' is not present on device"};
    }
    if (e->nbytes != nbytes) {
        // accelerator.cpp:419-427
        throw tbr::Error{TB_ERR_ARG, std::string(what) + ": buffer '" + (name ? name : "") +
                                         "' size does not match the device copy"};
    }
    return e;
}

int tb_accel_update_device(const void *host, size_t nbytes, const char *name) {
    TB_API_BEGIN
    tbr::require_device();
    std::lock_guard<std::mutex> lk(tbr::g_mu);
    tbr::TableEntry *e = need_entry(host, nbytes, name, "accel_update_device");
    TB_CUDA(cudaMemcpy(e->dev, host, nbytes, cudaMemcpyHostToDevice));
    TB_API_END
}

int tb_accel_update_host(void *host, size_t nbytes, const char *name) {
    TB_API_BEGIN
    tbr::require_device();
    std::lock_guard<std::mutex> lk(tbr::g_mu);
    tbr::TableEntry *e = need_entry(host, nbytes, name, "accel_update_host");
    TB_CUDA(cudaMemcpy(host, e->dev, nbytes, cudaMemcpyDeviceToHost));
    TB_API_END
}

int tb_accel_reset(const void *host, size_t nbytes, const char *name) {
    TB_API_BEGIN
    tbr::require_device();
    std::lock_guard<std::mutex> lk(tbr::g_mu);
    tbr::TableEntry *e = need_entry(host, nbytes, name, "accel_reset");
    TB_CUDA(cudaMemsetAsync(e->dev, 0, nbytes, 0));
    TB_API_END
}

int tb_accel_delete(const void *host, size_t nbytes, const char *name) {
    TB_API_BEGIN
    tbr::require_device();
    std::lock_guard<std::mutex> lk(tbr::g_mu);
    tbr::TableEntry *e = need_entry(host, nbytes, name, "accel_delete");
    TB_CUDA(cudaFree(e->dev));
    tbr::g_table_bytes -= e->nbytes;
    tbr::g_table.erase(host);
    TB_API_END
}

void *tb_accel_device_ptr(const void *host) {
    if (!tbr::enabled()) return nullptr;
    return tbr::table_lookup(host);
}

void tb_accel_dump(void) {
    std::lock_guard<std::mutex> lk(tbr::g_mu);
    for (auto &kv : tbr::g_table) {
        printf("toast_b200 accel table: host=%p dev=%p bytes=%zu name=%s\n", kv.first,
               kv.second.dev, kv.second.nbytes, kv.second.name.c_str());
    }
    fflush(stdout);
}

size_t tb_accel_bytes_in_use(void) { return tbr::g_table_bytes; }

} // extern "C"
