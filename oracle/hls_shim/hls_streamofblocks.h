/* Stand-in for Vitis-HLS "hls_streamofblocks.h" (PIPO ping-pong buffers).  In C-simulation the
 * DATAFLOW region runs its functions one after the other, so the producer queues every block
 * it writes and the consumer pops them in order.  TEST INFRASTRUCTURE ONLY. */
#ifndef SGRACE_SHIM_HLS_SOB_H
#define SGRACE_SHIM_HLS_SOB_H
#include <deque>
#include <stdlib.h>
#include <string.h>
namespace hls {
template <typename B> class stream_of_blocks {
public:
    std::deque<B *> q_;
    ~stream_of_blocks() { for (B *b : q_) free(b); }
};
template <typename B> class write_lock {
    stream_of_blocks<B> &s_; B *blk_;
public:
    explicit write_lock(stream_of_blocks<B> &s) : s_(s), blk_((B *)calloc(1, sizeof(B))) {}
    ~write_lock() { s_.q_.push_back(blk_); }
    operator B &() { return *blk_; }
    auto &operator[](size_t i) { return (*blk_)[i]; }
};
template <typename B> class read_lock {
    stream_of_blocks<B> &s_; B *blk_;
public:
    explicit read_lock(stream_of_blocks<B> &s) : s_(s), blk_(s.q_.front()) { s_.q_.pop_front(); }
    ~read_lock() { free(blk_); }
    operator B &() { return *blk_; }
    auto &operator[](size_t i) { return (*blk_)[i]; }
};
}  // namespace hls
#endif
