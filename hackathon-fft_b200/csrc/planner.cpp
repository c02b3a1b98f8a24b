#include "planner.hpp"

#include <algorithm>
#include <cstring>

namespace b200fft {

std::string& last_error() {
  static thread_local std::string s;
  return s;
}

int fail(int status, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error() = buf;
  return status;
}

std::atomic<uint64_t> g_launch_count{0};

// _div_by (_utils.mojo:125-129): how many times `base` divides x, by repeated division.
static uint64_t div_by(uint64_t x, uint64_t base) {
  uint64_t k = 0;
  while (x >= base && x % base == 0) {
    ++k;
    if (x == base) break;
    x /= base;
  }
  return k;
}

// _times_divisible_by (_utils.mojo:132-152). Power-of-two bases take the
// ctz(length) / log2(base) shortcut, which counts on `length` itself.
uint64_t times_divisible(uint64_t length, uint64_t base) {
  if (base < 2 || length == 0) return 0;
  if ((base & (base - 1)) == 0) return (uint64_t)__builtin_ctzll(length) / (uint64_t)__builtin_ctzll(base);
  return div_by(length, base);
}

// _build_ordered_bases (_utils.mojo:163-183)
std::vector<uint32_t> build_ordered_bases(uint64_t length, std::vector<uint32_t> bases) {
  std::sort(bases.begin(), bases.end());
  uint64_t prod = 1;
  for (uint32_t b : bases) prod *= b;
  if (prod == length) {
    std::reverse(bases.begin(), bases.end());
    return bases;
  }
  std::vector<uint32_t> out;
  uint64_t processed = 1;
  for (size_t k = bases.size(); k-- > 0;) {
    uint64_t times = times_divisible(length, bases[k]);
    for (uint64_t t = 0; t < times; ++t) {
      out.push_back(bases[k]);
      processed *= bases[k];
    }
    if (processed == length) break;
  }
  return out;
}

// asserts of _get_ordered_bases_processed_list (_utils.mojo:205-220)
bool ordered_bases_valid(uint64_t length, const std::vector<uint32_t>& ordered) {
  if (ordered.empty()) return false;
  uint64_t prod = 1;
  for (uint32_t b : ordered) {
    if (b < 2) return false;
    prod *= b;
    if (prod > length) return false;
  }
  return prod == length;
}

// _estimate_best_bases (fft.mojo:49-104)
std::vector<uint32_t> estimate_best_bases(uint64_t length, bool gpu_target) {
  const uint64_t max_radix = 32, block = 1024;
  if (gpu_target && length / max_radix <= block) {
    uint64_t lo = std::max<uint64_t>((length + block - 1) / block, 2);
    std::vector<uint32_t> pot;
    uint64_t processed = 1;
    for (uint64_t r = lo; r <= max_radix; ++r) {
      uint64_t times = times_divisible(length / processed, r);
      for (uint64_t t = 0; t < times; ++t) {
        pot.push_back((uint32_t)r);
        processed *= r;
      }
      if (processed == length) {
        std::reverse(pot.begin(), pot.end());
        return pot;
      }
    }
  }
  static const uint32_t primes[25] = {97, 89, 83, 79, 73, 71, 67, 61, 59, 53, 47, 43, 41,
                                      37, 31, 29, 23, 19, 17, 13, 11, 7,  5,  3,  2};
  std::vector<uint32_t> out;
  uint64_t processed = 1;
  for (uint32_t p : primes) {
    uint64_t times = times_divisible(length / processed, p);
    for (uint64_t t = 0; t < times; ++t) {
      out.push_back(p);
      processed *= p;
    }
    if (processed == length) break;
  }
  std::reverse(out.begin(), out.end());
  return out;
}

static size_t dtype_size(int dt) { return dt == B200FFT_U8 ? 1 : dt == B200FFT_F32 ? 4 : dt == B200FFT_F64 ? 8 : 0; }

int validate(const b200fft_desc* d, Problem* out) {
  if (!d || !out) return fail(B200FFT_ERR_INVALID_ARG, "null descriptor");
  // rank > 2 in the reference counts batch and complex dims: here rank = non-batch axes
  if (d->rank < 1 || d->rank > B200FFT_MAX_RANK)
    return fail(B200FFT_ERR_LAYOUT, "rank must be in 1..%d (got %d): layouts are (batches, dim_0[, dim_1...], 1|2)",
                B200FFT_MAX_RANK, d->rank);
  if (d->batch < 1) return fail(B200FFT_ERR_LAYOUT, "batch must be >= 1 (got %lld)", (long long)d->batch);
  if (d->in_components != 1 && d->in_components != 2)
    return fail(B200FFT_ERR_LAYOUT, "The last dimension of in_layout should be 1 or 2 (got %d)", d->in_components);
  if (dtype_size(d->in_dtype) == 0) return fail(B200FFT_ERR_INVALID_ARG, "unknown in_dtype %d", d->in_dtype);
  if (d->out_dtype != B200FFT_F32 && d->out_dtype != B200FFT_F64)
    return fail(B200FFT_ERR_LAYOUT, "out_dtype must be floating point (f32 or f64)");
  if (d->real_mode != B200FFT_REAL_FULL && d->real_mode != B200FFT_REAL_HALF)
    return fail(B200FFT_ERR_INVALID_ARG, "unknown real_mode %d", d->real_mode);

  Problem p;
  p.desc = *d;
  p.desc.bases = nullptr;
  p.desc.bases_count = nullptr;
  p.rank = d->rank;
  p.batch = d->batch;
  p.half = d->real_mode == B200FFT_REAL_HALF;
  p.in_elem = dtype_size(d->in_dtype);
  p.out_elem = dtype_size(d->out_dtype);
  const uint32_t mask = d->axis_mask ? d->axis_mask : ((1u << d->rank) - 1u);
  if (mask >> d->rank) return fail(B200FFT_ERR_INVALID_ARG, "axis_mask has bits beyond rank");

  const uint32_t* bp = d->bases;
  int64_t prod = 1;
  for (int a = 0; a < d->rank; ++a) {
    AxisSpec ax;
    ax.n = d->dims[a];
    ax.transformed = (mask >> a) & 1u;
    if (ax.n < 2 && !(ax.n == 1 && !ax.transformed))
      return fail(B200FFT_ERR_LAYOUT, "no inner dimension should be of size 1 (dims[%d] = %lld)", a, (long long)ax.n);
    if (ax.n > (int64_t)1 << 31) return fail(B200FFT_ERR_UNSUPPORTED, "dims[%d] too large", a);
    int cnt = (d->bases && d->bases_count) ? d->bases_count[a] : 0;
    if (cnt < 0) return fail(B200FFT_ERR_INVALID_ARG, "negative bases_count[%d]", a);
    if (cnt > 0) {
      ax.user.assign(bp, bp + cnt);
      bp += cnt;
    } else {
      ax.user = estimate_best_bases((uint64_t)ax.n, /*gpu_target=*/true);
    }
    if (ax.transformed) {
      for (uint32_t b : ax.user)
        if (b < 2) return fail(B200FFT_ERR_BASES, "Cannot do an fft with base %u (axis %d)", b, a);
      ax.ordered = build_ordered_bases((uint64_t)ax.n, ax.user);
      if (!ordered_bases_valid((uint64_t)ax.n, ax.ordered)) {
        std::string got;
        for (uint32_t b : ax.ordered) got += (got.empty() ? "" : ", ") + std::to_string(b);
        return fail(B200FFT_ERR_BASES,
                    "powers of the bases must multiply together to equal the sequence length. The builtin "
                    "algorithm was only able to produce: [%s] for the length: %lld (axis %d)",
                    got.c_str(), (long long)ax.n, a);
      }
    }
    prod *= ax.n;
    p.axes.push_back(std::move(ax));
  }

  if (p.half) {
    const AxisSpec& last = p.axes[d->rank - 1];
    if (!last.transformed) return fail(B200FFT_ERR_INVALID_ARG, "REAL_HALF needs the last axis transformed");
    // odd last axes are served by the generic kernel's Hermitian load / store (the fused R2C/C2R kernels need n = 2H)
    if (d->in_dtype == B200FFT_U8) return fail(B200FFT_ERR_UNSUPPORTED, "REAL_HALF takes floating-point input");
    const int64_t lead = prod / last.n, hc = last.n / 2 + 1;
    if (!d->inverse) {
      if (d->in_components != 1) return fail(B200FFT_ERR_LAYOUT, "forward REAL_HALF takes real input (in_components = 1)");
      p.in_scalars_per_batch = prod;
      p.out_scalars_per_batch = lead * hc * 2;
    } else {
      if (d->in_components != 2) return fail(B200FFT_ERR_LAYOUT, "inverse REAL_HALF takes the complex half spectrum");
      p.in_scalars_per_batch = lead * hc * 2;
      p.out_scalars_per_batch = prod;
    }
  } else {
    p.in_scalars_per_batch = prod * d->in_components;
    p.out_scalars_per_batch = prod * 2;
  }
  *out = std::move(p);
  return B200FFT_OK;
}

std::vector<NdSegment> build_schedule(int nphases, const SchedPhase* ph, long long batch) {
  std::vector<NdSegment> segs;
  std::vector<long long> prog(nphases, 0), total(nphases, 0);
  for (int p = 0; p < nphases; ++p) total[p] = ph[p].tiles_per_transform * batch;
  long long item = 0;
  while (true) {
    bool done = true;
    for (int p = 0; p < nphases; ++p) done = done && prog[p] == total[p];
    if (done) break;
    const std::vector<long long> before(prog);  // what earlier rounds handed out
    bool any = false;
    for (int p = 0; p < nphases; ++p) {
      long long limit = total[p];
      if (p > 0) {
        const long long groups_done =
            before[p - 1] == total[p - 1] ? (total[p] + ph[p].dep_div - 1) / ph[p].dep_div  // ragged last group
                                          : before[p - 1] / ph[p - 1].tiles_per_group;
        limit = std::min(limit, groups_done * ph[p].dep_div);
      }
      long long n = limit - prog[p];
      if (ph[p].quota > 0) n = std::min(n, ph[p].quota);
      if (n <= 0) continue;
      segs.push_back(NdSegment{p, 0, item, prog[p], n});
      prog[p] += n;
      item += n;
      any = true;
    }
    if (!any) break;  // cannot happen for consistent phase descriptions; avoids an endless loop on bad input
  }
  return segs;
}

}  // namespace b200fft
