// bsyn_prep.cu -- host-side data preparation in native code: the step before the hot path.
// Follows R/bayesynergy.R:129-154: unique non-zero dose levels per drug on the log10 scale, the position of every
// observation in the column-major (n2+1) x (n1+1) grid `f` (ii_obs, 1-based), replicates flattened column by
// column, missing values dropped, nrep / nmissing.  The reference matches rows to grid points with
// match(round(., 15)) after a 10^log10(x) round trip (R/bayesynergy.R:137); here every row is matched to the
// level it was deduplicated into (exact comparison of the original concentrations), which is the same map
// without the dependence on decimal rounding.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/bsyn.h"

namespace {
int levels(const double *x, int n, int col, std::vector<double> &lv)
{
    lv.resize(n);
    for (int i = 0; i < n; ++i) {
        const double v = x[2 * i + col];
        if (!(v >= 0.0) || std::isinf(v)) return -1;       // negative, NaN or infinite concentration
        lv[i] = v;
    }
    std::sort(lv.begin(), lv.end());
    lv.erase(std::unique(lv.begin(), lv.end()), lv.end());
    return 0;
}
}  // namespace

extern "C" int bsyn_prepare(const double *y, const double *x, int32_t n, int32_t nrep_in, double *x1_out, int32_t *n1_out,
                            double *x2_out, int32_t *n2_out, double *y_out, int32_t *ii_out, int32_t *N_out,
                            int32_t *nrep_out, int32_t *nmissing_out)
{
    if (!y || !x || n < 1 || nrep_in < 1 || !x1_out || !x2_out || !y_out || !ii_out || !n1_out || !n2_out || !N_out ||
        !nrep_out || !nmissing_out)
        return BSYN_ERR_ARG;
    std::vector<double> l1, l2;
    if (levels(x, n, 0, l1) || levels(x, n, 1, l2)) return BSYN_ERR_ARG;
    if (l1.empty() || l2.empty() || l1[0] != 0.0 || l2[0] != 0.0) return BSYN_ERR_DOMAIN;   // "Missing monotherapy data!"
    const int n1 = (int)l1.size() - 1, n2 = (int)l2.size() - 1;
    for (int j = 0; j < n1; ++j) x1_out[j] = std::log10(l1[j + 1]);
    for (int i = 0; i < n2; ++i) x2_out[i] = std::log10(l2[i + 1]);
    // grid position of every row: expand.grid(X1, X2) runs over X1 fastest in R but `f` is (n2+1) x (n1+1)
    // column-major with rows = drug-2 level, so ii = a (n2 + 1) + b + 1 for (drug-1 level a, drug-2 level b)
    std::vector<int32_t> ii(n);
    std::vector<int32_t> count((size_t)(n1 + 1) * (n2 + 1), 0);
    for (int i = 0; i < n; ++i) {
        const int a = (int)(std::lower_bound(l1.begin(), l1.end(), x[2 * i]) - l1.begin());
        const int b = (int)(std::lower_bound(l2.begin(), l2.end(), x[2 * i + 1]) - l2.begin());
        ii[i] = a * (n2 + 1) + b + 1;
        ++count[ii[i] - 1];
    }
    int nrep = nrep_in;
    if (nrep_in == 1) nrep = *std::max_element(count.begin(), count.end());   // replicates given as repeated rows
    int N = 0;
    for (int r = 0; r < nrep_in; ++r)
        for (int i = 0; i < n; ++i) {
            const double v = y[(size_t)r * n + i];
            if (std::isnan(v)) continue;
            y_out[N] = v;
            ii_out[N] = ii[i];
            ++N;
        }
    *n1_out = n1; *n2_out = n2; *N_out = N; *nrep_out = nrep;
    *nmissing_out = (n1 + n2 + n1 * n2 + 1) * nrep - N;
    return BSYN_OK;
}
