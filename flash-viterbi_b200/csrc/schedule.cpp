// schedule.cpp — the FLASH task tree as a pure function of (T, N).
//
// Reference behaviour restated (not copied): calc() (F:338-368) seeds a FIFO with either the N
// segments of the N-way pass (F:349-353, boundaries from F:129-136) or the single interval
// (0,T-1) (F:357-359); worker() (F:284-304) pops (L,R), solves mid=(L+R)>>1 and pushes (L,mid)
// and — when R > mid+1 — (mid+1,R), unless R <= L+1.  The queue order is breadth-first, so the
// levels built here, concatenated, are the reference's queue.
//   F: = /root/reference/src/FLASH_Viterbi_multithread.c
#include <pthread.h>

#include <algorithm>

#include "flashv_internal.h"

namespace flashv {

int pool_struct_bytes(int N)
{
    // sizeof(ThreadPool), F:36-43: mutex, condvar, N thread handles, three ints, 8-byte aligned.
    size_t raw = sizeof(pthread_mutex_t) + sizeof(pthread_cond_t) + sizeof(pthread_t) * (size_t)N + 3 * sizeof(int);
    return (int)((raw + 7u) & ~(size_t)7u);
}

bool build_schedule(int T, int N, Schedule *out)
{
    if (T < 2 || N < 1) return false;
    if (N > 2 && T == 2 * N) return false;  // the reference leaves Ans[] entries unset here
    Schedule s;
    s.T = T;
    s.N = N;
    s.first_pass = (N > 2 && T >= 2 * N);  // F:342
    std::vector<Task> level;
    if (s.first_pass) {
        const int span = T - 1;
        int gap = span / N, extra = span % N, at = 0;
        for (int t = 0; t + 1 < N; ++t) {  // F:129-136
            at += gap;
            if (extra) --extra, ++at;
            s.mids.push_back(at);
        }
        int lo = 0;
        for (int t = 0; t + 1 < N; ++t) {  // F:349-352
            level.push_back({lo, s.mids[t], (lo + s.mids[t]) >> 1});
            lo = s.mids[t] + 1;
        }
        level.push_back({lo, T - 1, (lo + T - 1) >> 1});  // F:353
        s.executed_steps = T - 1;
    } else {
        level.push_back({0, T - 1, (T - 1) >> 1});  // F:357-359
    }
    while (!level.empty()) {
        std::vector<Task> next;
        for (const Task &t : level) {
            s.fifo.push_back(t);
            s.executed_steps += t.R - t.L;
            if (t.R <= t.L + 1) continue;  // F:293
            next.push_back({t.L, t.mid, (t.L + t.mid) >> 1});  // F:300
            if (t.R > t.mid + 1) next.push_back({t.mid + 1, t.R, (t.mid + 1 + t.R) >> 1});  // F:301-302
        }
        std::stable_sort(level.begin(), level.end(),
                         [](const Task &a, const Task &b) { return (a.R - a.L) > (b.R - b.L); });
        s.levels.push_back(level);
        level.swap(next);
    }
    const size_t want = s.first_pass ? (size_t)(T - N) : (size_t)(T - 1);  // F:285: stops after queue index T-2
    if (s.fifo.size() != want) return false;
    *out = std::move(s);
    return true;
}

}  // namespace flashv

extern "C" int flashv_task_list(int T, int N, int *L, int *R, int *first_pass, int *mids)
{
    flashv::Schedule s;
    if (!flashv::build_schedule(T, N, &s)) {
        flashv::set_error("flashv_task_list: unsupported (T=%d, N=%d)", T, N);
        return FLASHV_ERR_ARG;
    }
    if (first_pass) *first_pass = s.first_pass ? 1 : 0;
    if (mids)
        for (size_t t = 0; t < s.mids.size(); ++t) mids[t] = s.mids[t];
    for (size_t q = 0; q < s.fifo.size(); ++q) {
        if (L) L[q] = s.fifo[q].L;
        if (R) R[q] = s.fifo[q].R;
    }
    return (int)s.fifo.size();
}

extern "C" long long flashv_executed_steps(int T, int N)
{
    flashv::Schedule s;
    if (!flashv::build_schedule(T, N, &s)) return FLASHV_ERR_ARG;
    return s.executed_steps;
}

extern "C" int flashv_memory_bytes(int K, int T, int N)
{
    // F:355 (the N-way pass's VLAs) vs F:364 (per-worker scratch), plus F:367.
    int mem = 0;
    if (N > 2 && T >= 2 * N)
        mem = (int)(sizeof(int) * (size_t)(N - 1) + sizeof(float) * 2u * (size_t)K +
                    sizeof(int) * 2u * (size_t)(N - 1) * (size_t)K);
    int per = N * (int)(2 * K * sizeof(float) + 2 * K * sizeof(int));
    if (per > mem) mem = per;
    return mem + flashv::pool_struct_bytes(N) + (int)sizeof(size_t);
}

extern "C" int flashv_bs_memory_bytes(int T, int N, int B)
{
    // S:564 vs S:573, plus S:576; sizeof(element) == 12 (S:51-56).
    int mem = 0;
    if (N > 2 && T >= 2 * N) mem = (int)(sizeof(int) * (size_t)(N - 1) + 12u * 2u * (size_t)(N - 1) * (size_t)(B + 1));
    int per = N * (int)(2 * (B + 1) * 12);
    if (per > mem) mem = per;
    return mem + flashv::pool_struct_bytes(N) + (int)sizeof(size_t);
}
