// c_api_common.hpp -- shared plumbing of the extern "C" layer (c_api.cpp, engine.cu).
#pragma once
#include <exception>
#include <new>
#include <string>

#include "../../include/mvf_b200.h"
#include "mvf_host.hpp"

struct mvf_space { mvf::VectorSpace vs; };

namespace mvf {

void set_last_error(const std::string& m, uint64_t first = 0, uint64_t second = 0);

// run `fn`, translate any exception into a status code + thread-local message
template <class Fn>
inline int guard(Fn&& fn) noexcept {
    try {
        fn();
        return MVF_OK;
    } catch (const MvfError& e) {
        set_last_error(e.what(), e.first, e.second);
        return (int)e.kind;
    } catch (const std::bad_alloc&) {
        set_last_error("I/O error: out of host memory");
        return MVF_ERR_IO;
    } catch (const std::exception& e) {
        set_last_error(std::string("Build error: ") + e.what());
        return MVF_ERR_BUILD;
    } catch (...) {
        set_last_error("Build error: unknown failure");
        return MVF_ERR_BUILD;
    }
}

// contiguous row shard of `rank`: ceil(n/world) rows per rank, the tail ranks may be short/empty
inline void shard_range(uint64_t n, int world, int rank, uint64_t* begin, uint64_t* count) {
    uint64_t per = (n + (uint64_t)world - 1) / (uint64_t)world;
    uint64_t b = per * (uint64_t)rank; if (b > n) b = n;
    uint64_t e = b + per; if (e > n) e = n;
    *begin = b; *count = e - b;
}

}  // namespace mvf
