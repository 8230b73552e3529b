// stand-in for boost::math::poisson_distribution<> with pdf / cdf (cdf(k) = Q(floor(k)+1, mean))
#pragma once
#include <cmath>
namespace boost { namespace math {
template <class RealType = double>
class poisson_distribution {
  RealType mean_;
 public:
  explicit poisson_distribution(RealType mean = 1) : mean_(mean) {}
  RealType mean() const { return mean_; }
};
typedef poisson_distribution<double> poisson;
template <class R> inline R pdf(const poisson_distribution<R>& d, const R& k) {
  if (k < 0) return 0;
  R kk = std::floor(k);
  if (d.mean() <= 0) return kk == 0 ? 1 : 0;
  return std::exp(-d.mean() + kk * std::log(d.mean()) - std::lgamma(kk + 1));
}
template <class R> inline R cdf(const poisson_distribution<R>& d, const R& k) {
  if (k < 0) return 0;
  if (d.mean() <= 0) return 1;
  long n = (long)std::floor(k);
  R s = 0, lm = std::log(d.mean());
  for (long j = 0; j <= n; j++) s += std::exp(-d.mean() + j * lm - std::lgamma((R)j + 1));
  return s > 1 ? 1 : s;
}
template <class R, class K> inline R pdf(const poisson_distribution<R>& d, const K& k) { return pdf(d, (R)k); }
template <class R, class K> inline R cdf(const poisson_distribution<R>& d, const K& k) { return cdf(d, (R)k); }
} }  // namespace boost::math
