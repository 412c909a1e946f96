// host_build.cpp -- see host_build.h
#include "host_build.h"

#include <algorithm>
#include <cstring>
#include <thread>

namespace gdx {

uint64_t storage_max(uint32_t storage) {
    switch (storage) {
    case GDX_I32: return 0x7fffffffull;
    case GDX_U32: return 0xffffffffull;
    default: return 0x7fffffffffffffffull;
    }
}

gdx_status concat_texts(const uint8_t *texts, const uint64_t *text_offsets, uint64_t num_texts,
                        const gdx_alphabet &alphabet, ConcatText &out, uint64_t *bad_text) {
    const uint32_t sigma = alphabet.num_dense_symbols;
    const uint64_t total = text_offsets[num_texts] - text_offsets[0];
    out.text.assign(total + num_texts, 0);
    out.sentinels.resize(num_texts);
    // sentinel positions first (construction/mod.rs:267-274), then the symbols are translated in parallel
    // slices of the concatenated output (the reference does this with rayon, construction/mod.rs:289-303)
    uint64_t w = 0;
    for (uint64_t t = 0; t < num_texts; ++t) {
        w += text_offsets[t + 1] - text_offsets[t];
        out.sentinels[t] = w++;
    }
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const unsigned nthreads = total < (1u << 22) ? 1 : hw;
    std::vector<std::vector<uint64_t>> freqs(nthreads, std::vector<uint64_t>(257, 0));
    std::vector<uint64_t> bad(nthreads, ~0ull);
    auto work = [&](unsigned tid) {
        // texts are split by their index range so that a slice never straddles a thread boundary mid-copy
        const uint64_t lo = total * tid / nthreads, hi = total * (tid + 1) / nthreads;  // input byte range
        // first text that contains input byte `lo`
        uint64_t t = std::upper_bound(text_offsets, text_offsets + num_texts + 1, text_offsets[0] + lo) - text_offsets - 1;
        std::vector<uint64_t> &freq = freqs[tid];
        for (uint64_t p = text_offsets[0] + lo; p < text_offsets[0] + hi;) {
            while (t + 1 <= num_texts && p >= text_offsets[t + 1]) ++t;  // skips empty texts
            const uint64_t end = std::min(text_offsets[0] + hi, text_offsets[t + 1]);
            uint8_t *dst = out.text.data() + (p - text_offsets[0]) + t;  // t sentinels precede text t
            for (; p < end; ++p) {
                const uint8_t d = alphabet.io_to_dense[texts[p]];
                if (d == 0) {  // alphabet.rs:195-198
                    bad[tid] = std::min(bad[tid], t);
                    return;
                }
                *dst++ = d;
                freq[d]++;
            }
        }
    };
    if (nthreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned i = 0; i < nthreads; ++i) th.emplace_back(work, i);
        for (auto &x : th) x.join();
    }
    uint64_t first_bad = ~0ull;
    for (uint64_t b : bad) first_bad = std::min(first_bad, b);
    if (first_bad != ~0ull) {
        if (bad_text) *bad_text = first_bad;
        return GDX_ERR_INVALID_SYMBOL;
    }
    std::vector<uint64_t> freq(257, 0);
    for (auto &f : freqs)
        for (int i = 0; i < 257; ++i) freq[i] += f[i];
    freq[0] = num_texts;  // construction/mod.rs:302
    out.count.assign(sigma + 1, 0);  // construction/mod.rs:318-336
    uint64_t sum = 0;
    for (uint32_t s = 0; s <= sigma; ++s) {
        out.count[s] = sum;
        sum += freq[s];
    }
    return GDX_OK;
}

// ---- SA-IS (Nong, Zhang, Chan 2009), induced sorting with explicit type bits ---------------------
namespace {

template <class I, class S>
struct Sais {
    const S *T;
    I *SA;
    I n, K;
    std::vector<bool> stype;
    std::vector<I> bkt;

    bool is_lms(I i) const { return i > 0 && stype[(size_t)i] && !stype[(size_t)i - 1]; }

    void buckets(bool end) {
        std::fill(bkt.begin(), bkt.end(), (I)0);
        for (I i = 0; i < n; ++i) bkt[(size_t)T[i]]++;
        I sum = 0;
        for (I c = 0; c < K; ++c) {
            sum += bkt[(size_t)c];
            bkt[(size_t)c] = end ? sum : sum - bkt[(size_t)c];
        }
    }
    void induce_l() {
        buckets(false);
        for (I i = 0; i < n; ++i) {
            const I j = SA[i] - 1;
            if (SA[i] > 0 && !stype[(size_t)j]) SA[bkt[(size_t)T[j]]++] = j;
        }
    }
    void induce_s() {
        buckets(true);
        for (I i = n - 1; i >= 0; --i) {
            const I j = SA[i] - 1;
            if (SA[i] > 0 && stype[(size_t)j]) SA[--bkt[(size_t)T[j]]] = j;
        }
    }

    void run() {
        stype.assign((size_t)n, false);
        bkt.assign((size_t)K, 0);
        stype[(size_t)n - 1] = true;
        for (I i = n - 2; i >= 0; --i)
            stype[(size_t)i] = T[i] < T[i + 1] || (T[i] == T[i + 1] && stype[(size_t)i + 1]);

        // stage 1: sort the LMS substrings
        buckets(true);
        std::fill(SA, SA + n, (I)-1);
        for (I i = 1; i < n; ++i)
            if (is_lms(i)) SA[--bkt[(size_t)T[i]]] = i;
        induce_l();
        induce_s();
        I n1 = 0;
        for (I i = 0; i < n; ++i)
            if (is_lms(SA[i])) SA[n1++] = SA[i];
        std::fill(SA + n1, SA + n, (I)-1);
        I name = 0, prev = -1;
        for (I i = 0; i < n1; ++i) {
            const I pos = SA[i];
            bool diff = false;
            for (I d = 0; d < n; ++d) {
                if (prev == -1 || T[pos + d] != T[prev + d] ||
                    stype[(size_t)(pos + d)] != stype[(size_t)(prev + d)]) {
                    diff = true;
                    break;
                } else if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) {
                    break;
                }
            }
            if (diff) {
                ++name;
                prev = pos;
            }
            SA[n1 + pos / 2] = name - 1;
        }
        for (I i = n - 1, j = n - 1; i >= n1; --i)
            if (SA[i] >= 0) SA[j--] = SA[i];

        // stage 2: order of the LMS suffixes
        I *SA1 = SA, *s1 = SA + n - n1;
        if (name < n1) {
            Sais<I, I> rec;
            rec.T = s1;
            rec.SA = SA1;
            rec.n = n1;
            rec.K = name;
            rec.run();
        } else {
            for (I i = 0; i < n1; ++i) SA1[s1[i]] = i;
        }

        // stage 3: induce the full order
        buckets(true);
        for (I i = 1, j = 0; i < n; ++i)
            if (is_lms(i)) s1[j++] = i;
        for (I i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
        std::fill(SA + n1, SA + n, (I)-1);
        for (I i = n1 - 1; i >= 0; --i) {
            const I j = SA[i];
            SA[i] = -1;
            SA[--bkt[(size_t)T[j]]] = j;
        }
        induce_l();
        induce_s();
    }
};

// suffix array of (text + 1) followed by a unique terminator 0 ("end of text < every symbol");
// full[0] is the terminator, full[1..] the suffix array of the text
template <class I, class S>
void sais_shifted_full(const uint8_t *text, uint64_t n, uint32_t sigma, std::vector<I> &full) {
    std::vector<S> t(n + 1);
    for (uint64_t i = 0; i < n; ++i) t[i] = (S)(text[i] + 1);
    t[n] = 0;
    full.assign(n + 1, 0);
    Sais<I, S> s;
    s.T = t.data();
    s.SA = full.data();
    s.n = (I)(n + 1);
    s.K = (I)(sigma + 1);
    s.run();
}

template <class S>
void sais_shifted(const uint8_t *text, uint64_t n, uint32_t sigma, std::vector<int64_t> &sa) {
    std::vector<int64_t> full;
    sais_shifted_full<int64_t, S>(text, n, sigma, full);
    sa.assign(full.begin() + 1, full.end());
}

template <class I>
void parts_from_sa(const uint8_t *text, uint64_t n, const I *sa, uint32_t sampling_rate, HostParts &out) {
    out.n = n;
    out.bwt.resize(n);
    out.samples.clear();
    out.samples.reserve(n / sampling_rate + 1);
    out.border_rows.clear();
    out.border_pos.clear();
    out.isa_samples.assign((n + sampling_rate - 1) / sampling_rate, 0);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t p = (uint64_t)sa[i];
        if (p % sampling_rate == 0) out.isa_samples[p / sampling_rate] = i;
        const uint8_t b = text[(p > 0 ? p : n) - 1];  // bwt.rs:96-105
        out.bwt[i] = b;
        if (b == 0) {  // bwt.rs:108-116
            out.border_rows.push_back(i);
            out.border_pos.push_back(p);
        }
        if (i % sampling_rate == 0) out.samples.push_back(p);  // sampled_suffix_array.rs:38-44
    }
}

}  // namespace

void suffix_array_sais(const uint8_t *text, uint64_t n, uint32_t sigma, std::vector<int64_t> &sa) {
    if (n == 0) {
        sa.clear();
        return;
    }
    if (sigma <= 255)
        sais_shifted<uint8_t>(text, n, sigma, sa);
    else
        sais_shifted<uint16_t>(text, n, sigma, sa);
}

void parts_from_suffix_array(const uint8_t *text, uint64_t n, const int64_t *sa,
                             uint32_t sampling_rate, HostParts &out) {
    parts_from_sa<int64_t>(text, n, sa, sampling_rate, out);
}

void host_parts_from_text(const uint8_t *text, uint64_t n, uint32_t sigma, uint32_t sampling_rate, HostParts &out) {
    if (n == 0) {
        out = HostParts();
        return;
    }
    if (n + 1 < (1ull << 31)) {
        std::vector<int32_t> full;
        if (sigma <= 255) sais_shifted_full<int32_t, uint8_t>(text, n, sigma, full);
        else sais_shifted_full<int32_t, uint16_t>(text, n, sigma, full);
        parts_from_sa<int32_t>(text, n, full.data() + 1, sampling_rate, out);
    } else {
        std::vector<int64_t> full;
        if (sigma <= 255) sais_shifted_full<int64_t, uint8_t>(text, n, sigma, full);
        else sais_shifted_full<int64_t, uint16_t>(text, n, sigma, full);
        parts_from_sa<int64_t>(text, n, full.data() + 1, sampling_rate, out);
    }
}

}  // namespace gdx
