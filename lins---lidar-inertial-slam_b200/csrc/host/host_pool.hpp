// host_pool.hpp — persistent host worker threads of a context (packing + staging of uploads).  PRODUCT code.
#ifndef LINS_HOST_POOL_HPP_
#define LINS_HOST_POOL_HPP_

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

// Persistent host workers of a context (packing + staging of uploads): created on first use, parked on a condition
// variable between uploads, so an upload does not pay for thread creation.
class HostPool {
 public:
  ~HostPool() {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  // run fn() on `n` threads in total (the caller is one of them) and wait for all of them
  void run(int n, const std::function<void()>& fn) {
    const int extra = std::max(0, n - 1);
    {
      std::lock_guard<std::mutex> lk(m_);
      while ((int)th_.size() < extra) th_.emplace_back([this] { loop(); });
      fn_ = &fn; want_ = extra; taken_ = 0; done_ = 0; ++gen_;
    }
    cv_.notify_all();
    fn();
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return done_ == want_; });
    fn_ = nullptr;
  }

 private:
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void()>* fn = nullptr;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || (gen_ != seen && taken_ < want_); });
        if (stop_) return;
        seen = gen_;
        ++taken_;
        fn = fn_;
      }
      (*fn)();
      {
        std::lock_guard<std::mutex> lk(m_);
        ++done_;
      }
      cv_done_.notify_one();
    }
  }
  std::mutex m_;
  std::condition_variable cv_, cv_done_;
  std::vector<std::thread> th_;
  const std::function<void()>* fn_ = nullptr;
  int want_ = 0, taken_ = 0, done_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
};

#endif
