// tb_runtime.cuh -- host-side runtime shared by the C-ABI translation units:
// error reporting, the host-pointer -> device-pointer table (the reference's OmpManager,
// accelerator.cpp:228-745), per-call staging of small host arrays, and the resolver that
// implements the three `mem` modes of include/toast_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/toast_b200.h"

namespace tbr {

void set_error(int code, const std::string &msg);
int last_code();

// Throwable carrying a TB_ERR_* code; caught by TB_API_BEGIN/END at the ABI boundary.
struct Error {
    int code;
    std::string msg;
};

#define TB_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            throw tbr::Error{TB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)}; \
        }                                                                                  \
    } while (0)

#define TB_REQUIRE(cond, msg)                                                              \
    do {                                                                                   \
        if (!(cond)) throw tbr::Error{TB_ERR_ARG, std::string(msg)};                       \
    } while (0)

#define TB_API_BEGIN try {
#define TB_API_END                                                                         \
    return TB_OK;                                                                          \
    }                                                                                      \
    catch (const tbr::Error &e) {                                                          \
        tbr::set_error(e.code, e.msg);                                                     \
        return e.code;                                                                     \
    }                                                                                      \
    catch (const std::exception &e) {                                                      \
        tbr::set_error(TB_ERR_ARG, e.what());                                              \
        return TB_ERR_ARG;                                                                 \
    }

// Makes sure a CUDA device is usable and selected; throws TB_ERR_NO_DEVICE otherwise.
void require_device();
void count_launch(int64_t n = 1);
int sm_count();

// ---- device memory table ------------------------------------------------------------------
struct TableEntry {
    void *dev;
    size_t nbytes;
    std::string name;
};
void *table_lookup(const void *host); // nullptr if absent

// ---- per-call resolver ---------------------------------------------------------------------
// Turns the pointers of one API call into device pointers according to `mem`, uploads the
// small host arrays, and on finish() copies outputs back / frees temporaries.
class Resolver {
  public:
    Resolver(int mem, void *stream);
    ~Resolver();

    cudaStream_t stream() const { return stream_; }
    int mem() const { return mem_; }

    // Large arrays.  `nbytes` is the full extent of the buffer.
    template <typename T> const T *in(const T *p, size_t count) {
        return (const T *)resolve((void *)p, count * sizeof(T), true, false);
    }
    template <typename T> T *inout(T *p, size_t count) {
        return (T *)resolve((void *)p, count * sizeof(T), true, true);
    }
    template <typename T> T *out(T *p, size_t count) {
        // outputs are only partially written (samples outside intervals keep their value),
        // so the previous content must travel too.
        return (T *)resolve((void *)p, count * sizeof(T), true, true);
    }
    // Small arrays: always host pointers, uploaded for this call.
    template <typename T> const T *small(const T *p, size_t count) {
        return (const T *)upload((const void *)p, count * sizeof(T));
    }
    // Small in/out host array (hit_submaps): uploaded now, downloaded at finish().
    template <typename T> T *small_inout(T *p, size_t count) {
        void *d = upload((const void *)p, count * sizeof(T));
        backs_.push_back(Back{d, (void *)p, count * sizeof(T), false});
        need_sync_ = true;
        return (T *)d;
    }
    // Device scratch owned by the call.
    void *scratch(size_t nbytes);

    // Copies outputs back, synchronises where the mode requires it, frees temporaries.
    void finish();

  private:
    struct Back {
        void *dev;
        void *host;
        size_t nbytes;
        bool free_dev;
    };
    void *resolve(void *p, size_t nbytes, bool copy_in, bool copy_out);
    void *upload(const void *p, size_t nbytes);
    int mem_;
    cudaStream_t stream_;
    bool need_sync_ = false;
    bool finished_ = false;
    std::vector<Back> backs_;
    std::vector<void *> temps_;
};

} // namespace tbr
