// Host path driver.  See path.h.  Every numerical step runs on the GPU through bess::Engine; this file only
// decides which sparsity level to evaluate next and applies the scalar criterion / de-normalisation formulas.
#include "path.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>
#include <random>

namespace bess {

// Metric::set_cv_train_test_mask (Metric.h:49-106): shuffle 0..n-1 with mt19937, folds are consecutive chunks of
// floor(n/K), the last fold takes the remainder.  The reference seeds from std::random_device (Metric.h:57-58);
// here the seed is an argument (parity on CV-chosen s needs the split pinned).
std::vector<int> cv_fold_ids(int n, int K, unsigned seed)
{
    std::vector<int> index_vec((size_t)n);
    std::iota(index_vec.begin(), index_vec.end(), 0);
    std::mt19937 g(seed);
    std::shuffle(index_vec.begin(), index_vec.end(), g);
    std::vector<int> fold((size_t)n, 0);
    const int group_size = n / K;
    for (int k = 0; k < K; k++) {
        const int b = k * group_size;
        const int e = (k == K - 1) ? n : b + group_size;
        for (int i = b; i < e; i++) fold[(size_t)index_vec[(size_t)i]] = k;
    }
    return fold;
}

namespace {

struct Eval {
    int T = 0, l = 0;
    std::vector<int> A;
    std::vector<double> bA;
    double coef0 = 0.0, train_loss = 0.0, ic = 0.0, lambda = 0.0;
};

struct Driver {
    Engine &eng;
    const BessArgs &a;
    int n, p, K;  // p: columns after screening; K: folds (0 when !is_cv)
    std::vector<int> all_chains, fold_chains;

    Driver(Engine &e, const BessArgs &args) : eng(e), a(args)
    {
        n = eng.n();
        p = (int)eng.p_model();
        K = a.is_cv ? a.K : 0;
        for (int c = 0; c <= K; c++) all_chains.push_back(c);
        for (int c = 1; c <= K; c++) fold_chains.push_back(c);
    }

    // Metric::ic without CV (Metric.h:205-229 gaussian: n*log(loss) + pen; :365-388 etc. GLMs: loss + pen).
    // The penalty uses the full sample size and the POST-screening column count.
    double ic_formula(double train_loss, int T) const
    {
        double pen;
        const double dn = (double)n, dp = (double)p;
        switch (a.ic_type) {
            case 1: pen = 2.0 * T; break;
            case 2: pen = std::log(dn) * T; break;
            case 3: pen = std::log(dp) * std::log(std::log(dn)) * T; break;
            case 4: pen = (std::log(dn) + 2 * std::log(dp)) * T; break;
            default: return 0.0;
        }
        if (a.model_type == 1) return dn * std::log(train_loss) + pen;
        return train_loss + pen;
    }

    static double mean_of(const std::vector<double> &v, size_t b, size_t e)
    {
        double s = 0.0;
        for (size_t i = b; i < e; i++) s += v[i];
        return s / (double)(e - b);
    }

    // One path step: full-data fit at level T (path.cpp:52-64) + train_loss + ic (path.cpp:71-72).  Under CV the K
    // fold fits of Metric::test_loss run in the same batch as the full fit (they never read its result; their
    // coef0_init is the one the path set BEFORE this step).
    // last_fold: also return the last fold's beta and ITS full-data loss (gs_path final sweep, path.cpp:314-319).
    Eval step(int T, Eval *last_fold = nullptr, double lambda = 0.0)
    {
        BatchResult br;
        std::vector<LossJob> jobs;
        jobs.push_back({0, 0, 0});
        for (int k = 0; k < K; k++) jobs.push_back({1 + k, 1, k});
        if (last_fold && K > 0) jobs.push_back({K, 0, 0});
        std::vector<double> v;
        eng.run_batch(T, all_chains, /*new_path_step=*/true, br, &jobs, &v, lambda);
        Eval e;
        e.T = T;
        e.lambda = lambda;
        e.l = br.l[0];
        e.A = br.A[0];
        e.bA = br.bA[0];
        e.coef0 = br.coef0[0];
        e.train_loss = v[0];
        e.ic = K > 0 ? mean_of(v, 1, 1 + (size_t)K) : ic_formula(e.train_loss, T);
        if (last_fold) {
            if (K > 0) {
                last_fold->T = T;
                last_fold->l = br.l[K];
                last_fold->A = br.A[K];
                last_fold->bA = br.bA[K];
                last_fold->coef0 = br.coef0[K];
                last_fold->train_loss = v[1 + (size_t)K];
                last_fold->ic = e.ic;
            } else {
                *last_fold = e;
            }
        }
        return e;
    }

    // A repeated metric->ic() on the same level (gs_path evaluates every fresh point twice): under CV
    // that is a second K-fold pass warm-started by the first; without CV the value is unchanged.
    double ic_again(const Eval &e)
    {
        if (K == 0) return e.ic;
        BatchResult br;
        std::vector<LossJob> jobs;
        for (int k = 0; k < K; k++) jobs.push_back({1 + k, 1, k});
        std::vector<double> v;
        eng.run_batch(e.T, fold_chains, /*new_path_step=*/false, br, &jobs, &v);
        return mean_of(v, 0, (size_t)K);
    }

    // path.cpp:76-110 / :330-343.  bA_out: de-normalised coefficients on the support e.A
    void denormalise(const Eval &e, std::vector<double> &bA_out, double &coef0) const
    {
        bA_out.assign(e.bA.begin(), e.bA.end());
        coef0 = e.coef0;
        if (!a.is_normal) return;
        const double sn = std::sqrt((double)n);
        double dot = 0.0;
        for (size_t i = 0; i < e.A.size(); i++) {
            const int j = e.A[i];
            const double b = sn * e.bA[i] / eng.x_norm_at(j);
            bA_out[i] = b;
            dot += b * eng.x_mean_at(j);
        }
        if (a.data_type == 1) coef0 = eng.y_mean() - dot;
        else if (a.data_type == 2) coef0 = e.coef0 - dot;
        // data_type 3: x_mean == 0, coef0 unchanged in both path functions
    }
};

void sequential_path(Driver &dr, BessResult &out, Eval &best)
{
    // path.cpp:48-74: for every sparsity level the lambda grid is walked zig-zag (:50), the warm start follows the walk.
    const BessArgs &a = dr.a;
    std::vector<double> lams = a.lambda_seq.empty() ? std::vector<double>{0.0} : a.lambda_seq;
    const int S = (int)a.sequence.size(), L = (int)lams.size();
    std::vector<Eval> evs((size_t)S * L);  // [lambda][s]
    std::vector<int> order;
    for (int i = 0; i < S; i++) {
        for (int q = 0; q < L; q++) {
            const int j = (i % 2 == 0) ? q : L - 1 - q;
            evs[(size_t)j * S + i] = dr.step(a.sequence[(size_t)i], nullptr, lams[(size_t)j]);
            order.push_back(j * S + i);
        }
    }
    // ic_sequence.minCoeff (path.cpp:113): Eigen visits the column-major matrix ic(s, lambda) column by column, i.e.
    // lambda outer / s inner, and keeps the first minimum
    size_t bi = 0;
    for (size_t i = 1; i < evs.size(); i++)
        if (evs[i].ic < evs[bi].ic) bi = i;
    best = evs[bi];
    for (int idx : order) {  // trace in evaluation order
        const Eval &e = evs[(size_t)idx];
        std::vector<double> b;
        double c0;
        dr.denormalise(e, b, c0);
        out.A_all.push_back(e.A);
        out.bA_all.push_back(std::move(b));
        out.coef0_all.push_back(c0);
        out.train_loss_all.push_back(e.train_loss);
        out.ic_all.push_back(e.ic);
        out.s_all.push_back(e.T);
        out.l_all.push_back(e.l);
        out.lambda_all.push_back(e.lambda);
    }
}

inline int c_round(double v) { return (int)std::round(v); }

// path.cpp:134-389, including its stateful quirks.
void gs_path(Driver &dr, BessResult &out, Eval &best)
{
    const BessArgs &a = dr.a;
    int Tmin = a.s_min, Tmax = a.s_max;
    int T1 = c_round(0.618 * Tmin + 0.382 * Tmax);
    int T2 = c_round(0.382 * Tmin + 0.618 * Tmax);
    double ic_seq[4] = {0, 0, 0, 0};
    double icT1, icT2;
    auto record = [&](const Eval &e) {
        out.s_all.push_back(e.T);
        out.ic_all.push_back(e.ic);
        out.train_loss_all.push_back(e.train_loss);
        out.l_all.push_back(e.l);
    };
    Eval e1 = dr.step(T1);  // :174-187 -- the very first point is scored once
    ic_seq[1] = e1.ic;
    icT1 = ic_seq[1];
    record(e1);
    Eval e2 = dr.step(T2);  // :189-210 -- ic() twice
    ic_seq[2] = e2.ic;
    icT2 = dr.ic_again(e2);
    record(e2);
    while (T1 != T2) {
        if (icT1 < icT2) {
            Tmax = T2;
            ic_seq[3] = ic_seq[2];
            T2 = T1;
            ic_seq[2] = ic_seq[1];
            icT2 = ic_seq[1];
            T1 = c_round(0.618 * Tmin + 0.382 * Tmax);
            Eval e = dr.step(T1);
            ic_seq[1] = e.ic;
            icT1 = dr.ic_again(e);
            record(e);
        } else {
            Tmin = T1;
            ic_seq[0] = ic_seq[1];
            T1 = T2;
            ic_seq[1] = ic_seq[2];
            icT1 = ic_seq[2];
            T2 = c_round(0.382 * Tmin + 0.618 * Tmax);
            Eval e = dr.step(T2);
            ic_seq[2] = e.ic;
            icT2 = dr.ic_again(e);
            record(e);
        }
    }
    double best_ic = DBL_MAX;
    bool have = false;
    for (int T = Tmin; T <= Tmax; T++) {
        Eval lf;
        Eval e = dr.step(T, &lf);
        if (e.ic < best_ic) {
            // algorithm->get_beta() is read AFTER ic(): under CV it is the last fold's fit, and train_loss is that
            // beta's full-data loss (path.cpp:314-319)
            best = lf;
            best.ic = e.ic;
            best_ic = e.ic;
            have = true;
            record(best);
        }
    }
    if (!have) throw EngineError{"gs_path: no finite criterion value in the final sweep"};
}

}  // namespace

void bess_run(const BessArgs &a, BessResult &out)
{
    // ---- argument validation (the reference validates in its R/Python front-ends and dereferences null otherwise,
    // SURVEY Q19; here bad arguments are reported)
    if (!a.x || !a.y || !a.weight) throw EngineError{"x, y and weight must be non-null"};
    if (a.n < 2 || a.p < 1) throw EngineError{"need n >= 2 and p >= 1"};
    if (!(a.algorithm_type == 1 || a.algorithm_type == 2 || a.algorithm_type == 3 || a.algorithm_type == 5))
        throw EngineError{"algorithm_type must be 1 (PDAS), 2, 3 or 5 (bess.cpp:93)"};
    if (a.model_type < 1 || a.model_type > 4) throw EngineError{"model_type must be 1..4"};
    if (a.data_type < 1 || a.data_type > 3) throw EngineError{"data_type must be 1..3"};
    if (a.path_type != 1 && (a.algorithm_type == 5 || a.algorithm_type == 3))
        throw EngineError{"pgs_path (bsrr, algorithm_type 3/5 with path_type 2) is outside this build's scope"};
    for (double l : a.lambda_seq) {
        if (!(l >= 0.0)) throw EngineError{"lambda_seq entries must be >= 0"};
        if (l != 0.0 && a.path_type != 1)
            throw EngineError{"lambda != 0 is supported on the sequential path only (gs_path ignores lambda, path.cpp:134)"};
    }
    if (!a.g_index.empty()) {
        if ((int)a.g_index.size() != a.p) throw EngineError{"group selection (gsize > 1) is outside this build's scope"};
        for (int j = 0; j < a.p; j++)
            if (a.g_index[(size_t)j] != j) throw EngineError{"g_index must be 0..p-1 (no groups)"};
    }
    if (a.is_cv && (a.K < 2 || a.K > MAXC - 1)) throw EngineError{"K (nfolds) must be in [2, 15]"};
    const bool shard = a.world > 1;
    const long long p_all = shard ? a.p_total : a.p;
    if (shard && a.x_on_device == false && a.x == nullptr) throw EngineError{"sharded fit: x shard is null"};
    for (int j : a.always_select)
        if (j < 0 || j >= p_all) throw EngineError{"always_select index out of range"};

    Engine eng(a.device);
    eng.set_profiling(a.profile);
    if (shard) eng.init_shard(a.world, a.rank, a.nccl_id, a.col_lo, a.p_total);
    eng.load(a.x, a.n, a.p, a.x_on_device, a.y, a.weight, a.model_type, /*borrow=*/a.is_screening);

    std::vector<int> always = a.always_select;
    std::sort(always.begin(), always.end());
    if (a.is_screening) {
        if (a.screening_size < 1 || a.screening_size > p_all) throw EngineError{"screening_size must be in [1, p]"};
        out.screening_A = eng.screen(a.screening_size, always);
        // screening.cpp:91-102: always_select -> positions inside the screened matrix
        for (int &j : always) {
            auto it = std::lower_bound(out.screening_A.begin(), out.screening_A.end(), j);
            if (it == out.screening_A.end() || *it != j) throw EngineError{"always_select column lost in screening"};
            j = (int)(it - out.screening_A.begin());
        }
    }
    eng.normalize(a.data_type, a.is_normal);

    const long long p = eng.p_model();
    int kcap;
    if (a.path_type == 1) {
        if (a.sequence.empty()) throw EngineError{"sequence (s.list) is empty"};
        kcap = *std::max_element(a.sequence.begin(), a.sequence.end());
        if (*std::min_element(a.sequence.begin(), a.sequence.end()) < 1) throw EngineError{"s.list entries must be >= 1"};
    } else {
        if (a.s_min < 1 || a.s_max < a.s_min) throw EngineError{"need 1 <= s_min <= s_max"};
        kcap = a.s_max;
    }
    if (kcap > p) throw EngineError{"sparsity level exceeds the number of (screened) columns"};

    std::vector<int> folds;
    if (a.is_cv) {
        if (a.fold_of_row) folds.assign(a.fold_of_row, a.fold_of_row + a.n);
        else folds = cv_fold_ids(a.n, a.K, a.cv_seed);
    }
    eng.setup_chains(a.is_cv ? a.K : 0, folds.data(), kcap, a.max_iter, a.is_warm_start, always);

    Driver dr(eng, a);
    Eval best;
    if (a.path_type == 1) sequential_path(dr, out, best);
    else gs_path(dr, out, best);

    std::vector<double> bA;
    double coef0;
    dr.denormalise(best, bA, coef0);
    out.coef0 = coef0;
    out.train_loss = best.train_loss;
    out.ic = best.ic;
    out.lambda = best.lambda;
    out.chosen_s = best.T;
    // scatter to the ORIGINAL column numbering (un-screen: bess.cpp:186-209)
    out.beta.assign((size_t)p_all, 0.0);
    auto orig = [&](int j) { return a.is_screening ? out.screening_A[(size_t)j] : j; };
    for (size_t i = 0; i < best.A.size(); i++) out.beta[(size_t)orig(best.A[i])] = bA[i];
    for (auto &A : out.A_all)
        for (int &j : A) j = orig(j);
    out.stats = eng.stats();
    eng.profile(out.prof_ms, out.prof_n);
    out.sweep_splits = eng.sweep_splits();
}

}  // namespace bess
