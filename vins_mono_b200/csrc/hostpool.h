// Minimal fork-join pool for the per-member host phases of a batch (the members' bookkeeping is independent).
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace vb {

class HostPool {
  public:
    explicit HostPool(int workers) {
        for (int i = 0; i < workers; i++) th_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // fn(i) for i in [0, n); the caller takes part; returns when all are done
    void run(int n, const std::function<void(int)>& fn) {
        if (th_.empty() || n <= 1) {
            for (int i = 0; i < n; i++) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            pending_ = (int)th_.size();
            gen_++;
        }
        cv_.notify_all();
        for (int i; (i = next_.fetch_add(1)) < n;) fn(i);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
    static int default_workers(int members) {
        const int hw = (int)std::thread::hardware_concurrency();
        return std::max(0, std::min(std::min(members - 1, hw - 2), 23));
    }

  private:
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            int n;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                fn = fn_;
                n = n_;
            }
            for (int i; (i = next_.fetch_add(1)) < n;) (*fn)(i);
            std::lock_guard<std::mutex> lk(m_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, pending_ = 0;
    std::atomic<int> next_{0};
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

}  // namespace vb
