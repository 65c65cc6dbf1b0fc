/* Stand-in for Vitis-HLS "hls_stream.h": an unbounded FIFO, which is what hls::stream is in
 * C-simulation.  TEST INFRASTRUCTURE ONLY; written for this repo. */
#ifndef SGRACE_SHIM_HLS_STREAM_H
#define SGRACE_SHIM_HLS_STREAM_H
#include <deque>
namespace hls {
template <typename T> class stream {
    std::deque<T> q_;
public:
    stream() {}
    stream(const char *) {}
    void write(const T &v) { q_.push_back(v); }
    T read() { T v = q_.front(); q_.pop_front(); return v; }
    void operator<<(const T &v) { write(v); }
    void operator>>(T &v) { v = read(); }
    bool read_nb(T &v) { if (q_.empty()) return false; v = read(); return true; }
    bool write_nb(const T &v) { write(v); return true; }
    bool empty() const { return q_.empty(); }
    bool full() const { return false; }
    size_t size() const { return q_.size(); }
};
}  // namespace hls
#endif
