// A process-wide pool of host threads for large memcpy()s (staging pageable columns through pinned slots,
// aero_b200/csrc/abi.cu parallel_memcpy).  Header-only so that the CPU test suite can stress it without CUDA.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace aero {
namespace host {

class CopyPool {
  public:
    struct Chunk {
        uint8_t *dst;
        const uint8_t *src;
        size_t len;
    };
    explicit CopyPool(unsigned workers) {
        for (unsigned i = 0; i < workers; i++) std::thread([this] { work(); }).detach();
    }
    // Heap-allocate and never free: the detached workers sleep on this object's condition variable, and
    // destroying a condition variable with waiters blocks (glibc) or is undefined.
    ~CopyPool() = delete;
    // copies all chunks; returns when every byte has been written.  One batch at a time.
    void run(std::vector<Chunk> &&chunks) {
        std::lock_guard<std::mutex> one(run_mu_);
        std::unique_lock<std::mutex> lk(mu_);
        chunks_ = std::move(chunks);
        next_ = 0;
        pending_ = chunks_.size();
        cv_work_.notify_all();
        while (next_ < chunks_.size()) {  // the caller works as well
            const Chunk c = chunks_[next_++];
            lk.unlock();
            memcpy(c.dst, c.src, c.len);
            lk.lock();
            pending_--;
        }
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        chunks_.clear();
        next_ = 0;
    }

  private:
    void work() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_work_.wait(lk, [&] { return next_ < chunks_.size(); });
            const Chunk c = chunks_[next_++];
            lk.unlock();
            memcpy(c.dst, c.src, c.len);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }
    std::mutex run_mu_, mu_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<Chunk> chunks_;
    size_t next_ = 0, pending_ = 0;
};

}  // namespace host
}  // namespace aero
