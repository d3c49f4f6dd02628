// Reflector log between stage B (QR iteration on H) and stage C (replay on the Schur vectors).
//
// Nothing on the serial chain of a Francis sweep ever reads Z: the Schur vectors only need the reflectors, in order.
// Stage B therefore writes every transformation it applies from the right (reflectors, the unit-modulus column
// scalings of src/GenericSchur.jl:461-500, the Givens rotation of the real 2x2 standardisation :674-687) into a log in
// global memory, and stage C (zreplay.cuh) applies the log to Z with the whole of Z in shared memory and one thread
// per ROW of Z — rows are independent, so stage C has no serial chain at all.
//
// Layout.  A record is four reals (32 B for Float64 / ComplexF64).  Records live in pages of LOG_PAGE_REC records
// taken from a pool with an atomic bump allocator; each matrix has a row in the page table:
//     row[0] = number of records, row[1] = status (0 ok, 1 overflow), row[2 + i] = id of the matrix's i-th page.
// Record stream of one matrix (k is the 1-based column index the reference uses):
//     header  {int op, int k, int count, int k2 | x, y}    (the ints overlay the first 16 bytes, x, y = a[2], a[3])
//       LOG_REFL   (complex): `count` payload records follow, reflector i acts on columns k+i, k+i+1:
//                             payload = {tau1.re, tau1.im, v2.re, v2.im}              (src/GenericSchur.jl:455-459)
//       LOG_SCALE  (complex): columns k..k2 are multiplied by x + iy; no payload      (:461-482, :486-500)
//       LOG_REFL3  (real):    `count` payload records {tau1, v2, v3, 0}, columns k+i .. k+i+2     (:920-925)
//       LOG_REFL2  (real):    one payload record {tau1, v2, 0, 0}, columns k, k+1                 (:940-945)
//       LOG_GIVENS (real):    columns k, k+1 rotated by (cs, sn) = (x, y); no payload             (:687)
// A matrix whose log does not fit (its per-matrix page budget or the pool is exhausted: only pathological inputs that
// iterate several times longer than random matrices) is flagged; stage B leaves its H untouched in global memory,
// stage C skips it, and the fused kernel (fastqr.cuh, Z streamed through L2) redoes it from the redo list.
#pragma once
#include "launch.h"
#include "scalar.cuh"

namespace gs {

enum { LOG_REFL = 1, LOG_SCALE = 2, LOG_REFL3 = 3, LOG_REFL2 = 4, LOG_GIVENS = 5 };

GS_DEV void stg_2f64_if(void* a, double x, double y, bool p) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %3, 0; @q st.global.v2.f64 [%0], {%1, %2}; }" ::"l"(a), "d"(x), "d"(y),
                 "r"((int)p)
                 : "memory");
}
GS_DEV void stg_4i_if(void* a, int x, int y, int z, int w, bool p) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %5, 0; @q st.global.v4.b32 [%0], {%1, %2, %3, %4}; }" ::"l"(a), "r"(x),
                 "r"(y), "r"(z), "r"(w), "r"((int)p)
                 : "memory");
}

// Producer side.  One "slot" of a warp (all 32 lanes, or an aligned group of 16 / 8 lanes when a warp works on several
// matrices at once) writes the log of one matrix: the state is uniform across the slot's lanes, the leader lane stores.
template <class R> struct LogWriter {
    static constexpr int REC = 4 * (int)sizeof(R);
    static constexpr int PAGE_BYTES = LOG_PAGE_REC * REC;
    unsigned char* pool;
    unsigned* next;
    unsigned npages;
    int* row;
    int maxp;
    unsigned char* cur;
    int left, nrec, npg;
    unsigned mask;     // lanes of the slot
    int leader;        // first lane of the slot
    bool lead;         // this lane is the leader
    bool on, ovf;

    GS_DEV void init(const BatchedParams& p, long long b, int lane, bool enable, unsigned slot_mask = 0xffffffffu,
                     int leader_lane = 0) {
        pool = p.log_pool;
        next = p.log_next;
        npages = p.log_pages;
        maxp = p.log_maxp;
        row = p.log_table ? p.log_table + b * (long long)(2 + p.log_maxp) : nullptr;   // a disabled log still reports "no records"
        cur = nullptr;
        left = 0;
        nrec = 0;
        npg = 0;
        mask = slot_mask;
        leader = leader_lane;
        lead = lane == leader_lane;
        on = enable && p.log_pool != nullptr;
        ovf = false;
    }
    GS_DEV void new_page() {
        if (npg >= maxp) {
            ovf = true;
            return;
        }
        unsigned pg = 0;
        if (lead) pg = atomicAdd(next, 1u);
        pg = __shfl_sync(mask, pg, leader);
        if (pg >= npages) {
            ovf = true;
            return;
        }
        if (lead) row[2 + npg] = (int)pg;
        npg += 1;
        cur = pool + (size_t)pg * PAGE_BYTES;
        left = LOG_PAGE_REC;
    }
    // the caller stores the record at `slot()` (leader lane) and then calls `advance()`
    GS_DEV bool reserve() {   // returns true when the record may be stored
        if (!on || ovf) return false;
        if (left == 0) new_page();
        return !ovf;
    }
    GS_DEV unsigned char* slot() const { return cur; }
    GS_DEV void advance() {
        cur += REC;
        left -= 1;
        nrec += 1;
    }
    GS_DEV void put_hdr(int op, int k, int count, int k2, const R& x, const R& y) {
        if (!reserve()) return;
        if (lead) {
            int* h = reinterpret_cast<int*>(cur);
            h[0] = op;
            h[1] = k;
            h[2] = count;
            h[3] = k2;
            R* a = reinterpret_cast<R*>(cur);
            a[2] = x;
            a[3] = y;
        }
        advance();
    }
    GS_DEV void put4(const R& a0, const R& a1, const R& a2, const R& a3) {
        if (!reserve()) return;
        if (lead) {
            R* a = reinterpret_cast<R*>(cur);
            a[0] = a0;
            a[1] = a1;
            a[2] = a2;
            a[3] = a3;
        }
        advance();
    }
    GS_DEV void finish() {
        if (row && lead) {
            row[0] = on ? nrec : 0;
            row[1] = ovf ? 1 : 0;
        }
    }
};

// number of log records a matrix of order n is expected to produce (random dense input), used to size the pool;
// measured: complex single shift ~1.7 n^2 reflectors + ~3.3 n sweep headers, real double shift ~1.0 n^2 + ~1.9 n
inline long long log_expected_records(bool cplx, int n) {
    const double n2 = (double)n * n;
    return (long long)((cplx ? 1.8 : 1.15) * n2 + 12.0 * n + 64.0);
}

}  // namespace gs
