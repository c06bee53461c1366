// TEST INFRASTRUCTURE ONLY.  Thin extern "C" driver around the *real* reference
// classes (included from /root/reference/src via -I; no reference source is
// copied).  Built into oracle/_ref/libbess_ref.so by oracle/Makefile together
// with the reference's own translation units.
//
// Entry points:
//   ref_pywrap_bess   -> /root/reference/src/bess.cpp:218 pywrap_bess (C++ linkage there)
//   ref_cv_fold_ids   -> /root/reference/src/Metric.h:49 Metric::set_cv_train_test_mask
//   ref_single_fit    -> /root/reference/src/Algorithm.h:113 Algorithm::fit (one PDAS fit)
//   ref_seq_trace     -> re-drives /root/reference/src/path.cpp:48-74 with the reference's own
//                        fit()/train_loss()/ic() and records every level (what the R build
//                        returns as beta_all / ic_all, path.cpp:116-123)
//   ref_bess_lambda   -> /root/reference/src/bess.cpp:37 bessCpp, result list incl. "lambda" (sequential_path with a
//                        lambda grid, path.cpp:25-132; pgs_path Powell search, path.cpp:1138-1309)
//   ref_pgs_trace     -> /root/reference/src/path.cpp:1138 pgs_path driven directly, every Algorithm::fit() logged
#include <Eigen/Eigen>
#include "List.h"
#include "Data.h"
#include "Algorithm.h"
#include "Metric.h"
#include "path.h"
#include "utilities.h"
#include "screening.h"
#include "bess.h"
#include <cstring>

static Algorithm *make_algorithm(Data &data, int model_type, int algorithm_type, int max_iter)
{
    // mirrors the dispatch in /root/reference/src/bess.cpp:93-112
    if (model_type == 1) { data.add_weight(); return new GroupPdasLm(data, algorithm_type, max_iter); }
    if (model_type == 2) return new GroupPdasLogistic(data, algorithm_type, max_iter);
    if (model_type == 3) return new GroupPdasPoisson(data, algorithm_type, max_iter);
    return new GroupPdasCox(data, algorithm_type, max_iter);
}
static Metric *make_metric(int model_type, int ic_type, bool is_cv, int K)
{
    if (model_type == 1) return new LmMetric(ic_type, is_cv, K);
    if (model_type == 2) return new LogisticMetric(ic_type, is_cv, K);
    if (model_type == 3) return new PoissonMetric(ic_type, is_cv, K);
    return new CoxMetric(ic_type, is_cv, K);
}

extern "C" {

void ref_pywrap_bess(double *x, int x_row, int x_col, double *y, int y_len, int data_type, double *weight, int weight_len,
                     bool is_normal, int algorithm_type, int model_type, int max_iter, int exchange_num,
                     int path_type, bool is_warm_start, int ic_type, bool is_cv, int K,
                     int *gindex, int gindex_len, double *state, int state_len, int *sequence, int sequence_len,
                     double *lambda_sequence, int lambda_sequence_len, int s_min, int s_max, int K_max, double epsilon,
                     double lambda_min, double lambda_max, int n_lambda, bool is_screening, int screening_size, int powell_path,
                     int *always_select, int always_select_len, double tao,
                     double *beta_out, int beta_out_len, double *coef0_out, int coef0_out_len, double *train_loss_out,
                     int train_loss_out_len, double *ic_out, int ic_out_len, double *nullloss_out, double *aic_out,
                     int aic_out_len, double *bic_out, int bic_out_len, double *gic_out, int gic_out_len, int *A_out,
                     int A_out_len, int *l_out)
{
    pywrap_bess(x, x_row, x_col, y, y_len, data_type, weight, weight_len, is_normal, algorithm_type, model_type, max_iter,
                exchange_num, path_type, is_warm_start, ic_type, is_cv, K, gindex, gindex_len, state, state_len, sequence,
                sequence_len, lambda_sequence, lambda_sequence_len, s_min, s_max, K_max, epsilon, lambda_min, lambda_max,
                n_lambda, is_screening, screening_size, powell_path, always_select, always_select_len, tao, beta_out,
                beta_out_len, coef0_out, coef0_out_len, train_loss_out, train_loss_out_len, ic_out, ic_out_len, nullloss_out,
                aic_out, aic_out_len, bic_out, bic_out_len, gic_out, gic_out_len, A_out, A_out_len, l_out);
}

// fold_of_row[i] = index of the CV fold whose *test* set holds row i.
void ref_cv_fold_ids(int n, int K, int *fold_of_row)
{
    LmMetric metric(1, true, K);
    metric.set_cv_train_test_mask(n);
    for (int k = 0; k < K; k++)
        for (int i = 0; i < metric.test_mask_list[k].size(); i++)
            fold_of_row[metric.test_mask_list[k](i)] = k;
}

// screening only: returns screening_A (ascending original indices)
void ref_screening(double *x, int n, int p, double *y, double *weight, int model_type, int screening_size, int *screening_A_out)
{
    Eigen::MatrixXd X = Pointer2MatrixXd(x, n, p);
    Eigen::VectorXd Y = Pointer2VectorXd(y, n);
    Eigen::VectorXd W = Pointer2VectorXd(weight, n);
    Eigen::VectorXi g_index = Eigen::VectorXi::LinSpaced(p, 0, p - 1);
    Eigen::VectorXi always_select(0);
    Eigen::VectorXi A = screening(X, Y, W, model_type, screening_size, g_index, always_select);
    for (int i = 0; i < screening_size; i++) screening_A_out[i] = A(i);
}

// One Algorithm::fit() on (optionally) a row subset, in the reference's *normalised* coordinates.
// train_mask: ascending row indices (length train_n).  beta_out is length p (normalised scale).
void ref_single_fit(double *x, int n, int p, double *y, double *weight, int data_type, bool is_normal, int model_type,
                    int max_iter, int T0, int *train_mask, int train_n, double *beta_init, double coef0_init,
                    double *beta_out, double *coef0_out, int *l_out)
{
    Eigen::MatrixXd X = Pointer2MatrixXd(x, n, p);
    Eigen::VectorXd Y = Pointer2VectorXd(y, n);
    Eigen::VectorXd W = Pointer2VectorXd(weight, n);
    Eigen::VectorXi g_index = Eigen::VectorXi::LinSpaced(p, 0, p - 1);
    Data data(X, Y, data_type, W, is_normal, g_index);
    Algorithm *alg = make_algorithm(data, model_type, 1, max_iter);
    alg->always_select = Eigen::VectorXi(0);
    Eigen::VectorXi mask = Pointer2VectorXi(train_mask, train_n);
    alg->update_train_mask(mask);
    alg->update_sparsity_level(T0);
    alg->update_beta_init(Pointer2VectorXd(beta_init, p));
    alg->update_coef0_init(coef0_init);
    if (model_type == 1) {
        Eigen::MatrixXd tx(train_n, p);
        for (int i = 0; i < train_n; i++) tx.row(i) = alg->data.x.row(mask(i));
        alg->update_group_XTX(group_XTX(tx, g_index, alg->data.g_size, train_n, p, p, 1));
    }
    alg->fit();
    Eigen::VectorXd b = alg->get_beta();
    for (int j = 0; j < p; j++) beta_out[j] = b(j);
    *coef0_out = alg->get_coef0();
    *l_out = alg->get_l();
    delete alg;
}

// Sequential path, every level recorded (normalised-scale beta, as inside path.cpp before :76).
// beta_all: [sequence_len][p], coef0_all/loss_all/ic_all/l_all: [sequence_len].
void ref_seq_trace(double *x, int n, int p, double *y, double *weight, int data_type, bool is_normal, int model_type,
                   int max_iter, bool is_warm_start, int ic_type, bool is_cv, int K, int *sequence, int sequence_len,
                   double *beta_all, double *coef0_all, double *loss_all, double *ic_all, int *l_all)
{
    Eigen::MatrixXd X = Pointer2MatrixXd(x, n, p);
    Eigen::VectorXd Y = Pointer2VectorXd(y, n);
    Eigen::VectorXd W = Pointer2VectorXd(weight, n);
    Eigen::VectorXi g_index = Eigen::VectorXi::LinSpaced(p, 0, p - 1);
    Data data(X, Y, data_type, W, is_normal, g_index);
    Algorithm *alg = make_algorithm(data, model_type, 1, max_iter);
    alg->set_warm_start(is_warm_start);
    alg->always_select = Eigen::VectorXi(0);
    Metric *metric = make_metric(model_type, ic_type, is_cv, K);
    if (is_cv) {
        metric->set_cv_train_test_mask(data.get_n());
        metric->set_cv_initial_model_param(K, data.get_p());
        if (model_type == 1) metric->cal_cv_group_XTX(data);
    }
    Eigen::VectorXi full_mask = Eigen::VectorXi::LinSpaced(n, 0, n - 1);
    std::vector<Eigen::MatrixXd> full_xtx = group_XTX(data.x, data.g_index, data.g_size, data.n, data.p, data.g_num, alg->model_type);
    Eigen::VectorXd beta_init = Eigen::VectorXd::Zero(p);
    double coef0_init = 0.0;
    for (int i = 0; i < sequence_len; i++) {
        alg->update_train_mask(full_mask);
        alg->update_sparsity_level(sequence[i]);
        alg->update_lambda_level(0.0);
        alg->update_beta_init(beta_init);
        alg->update_coef0_init(coef0_init);
        alg->update_group_XTX(full_xtx);
        alg->fit();
        if (alg->warm_start) { beta_init = alg->get_beta(); coef0_init = alg->get_coef0(); }
        Eigen::VectorXd b = alg->get_beta();
        for (int j = 0; j < p; j++) beta_all[(size_t)i * p + j] = b(j);
        coef0_all[i] = alg->get_coef0();
        l_all[i] = alg->get_l();
        loss_all[i] = metric->train_loss(alg, data);
        ic_all[i] = metric->ic(alg, data);
    }
    delete alg;
    delete metric;
}

// bessCpp (/root/reference/src/bess.cpp:37) with the "lambda" key of the result list surfaced as well (pywrap_bess
// drops it, bess.cpp:270-280): the ridge level chosen by sequential_path (path.cpp:127) / pgs_path (path.cpp:1290).
void ref_bess_lambda(double *x, int n, int p, double *y, int data_type, double *weight, bool is_normal, int algorithm_type,
                     int model_type, int max_iter, int path_type, bool is_warm_start, int ic_type, bool is_cv, int K,
                     int *sequence, int sequence_len, double *lambda_sequence, int lambda_sequence_len, int s_min,
                     int s_max, double lambda_min, double lambda_max, int n_lambda, bool is_screening, int screening_size,
                     int powell_path, int *always_select, int always_select_len, int *gindex, int gindex_len,
                     double *beta_out, double *coef0_out, double *train_loss_out, double *ic_out, double *lambda_out)
{
    Eigen::VectorXd state = Eigen::VectorXd::Zero(1);
    Eigen::VectorXi g_index = Pointer2VectorXi(gindex, gindex_len);  // first column of every group (bess.R:540)
    List res = bessCpp(Pointer2MatrixXd(x, n, p), Pointer2VectorXd(y, n), data_type, Pointer2VectorXd(weight, n), is_normal,
                       algorithm_type, model_type, max_iter, 2, path_type, is_warm_start, ic_type, is_cv, K, state,
                       Pointer2VectorXi(sequence, sequence_len), Pointer2VectorXd(lambda_sequence, lambda_sequence_len),
                       s_min, s_max, 10, 10.0, lambda_min, lambda_max, n_lambda, is_screening, screening_size, powell_path,
                       g_index, Pointer2VectorXi(always_select, always_select_len), 1.1);
    Eigen::VectorXd beta;
    double lam = 0.0;
    res.get_value_by_name("beta", beta);
    res.get_value_by_name("coef0", *coef0_out);
    res.get_value_by_name("train_loss", *train_loss_out);
    res.get_value_by_name("ic", *ic_out);
    res.get_value_by_name("lambda", lam);
    *lambda_out = lam;
    for (int j = 0; j < p; j++) beta_out[j] = beta(j);
}

// ---- pgs_path with every Algorithm::fit() logged.  The reference classes are used as they are; a derived class only
// overrides the virtual get_A (Algorithm.h:228) to note (sparsity_level, lambda_level, #train rows, coef0_init,
// nnz(beta_init)) on the first PDAS iteration of each fit (Algorithm::l == 1, Algorithm.h:151) before delegating.
}  // extern "C"

struct FitLog {
    std::vector<double> rec;  // 5 doubles per fit
};
template <class Base>
struct Traced : Base {
    FitLog *log;
    Traced(Data &data, int algorithm_type, int max_iter, FitLog *lg) : Base(data, algorithm_type, max_iter), log(lg) {}
    void get_A(Eigen::MatrixXd X, Eigen::VectorXd y, Eigen::VectorXd beta, double coef0, int T0, Eigen::VectorXd weights,
               Eigen::VectorXi index, Eigen::VectorXi gsize, int N, Eigen::VectorXi &A_out)
    {
        if (this->l == 1) {
            int nnz = 0;
            for (int j = 0; j < this->beta_init.size(); j++) nnz += this->beta_init(j) != 0.0;
            const double r[5] = {(double)this->sparsity_level, this->lambda_level, (double)X.rows(), this->coef0_init, (double)nnz};
            log->rec.insert(log->rec.end(), r, r + 5);
        }
        Base::get_A(X, y, beta, coef0, T0, weights, index, gsize, N, A_out);
    }
};

extern "C" {

// returns the number of fits; rec_out receives min(nfits, max_rec) records of 5 doubles
int ref_pgs_trace(double *x, int n, int p, double *y, double *weight, int data_type, bool is_normal, int algorithm_type,
                  int model_type, int max_iter, bool is_warm_start, int ic_type, bool is_cv, int K, int s_min, int s_max,
                  double lambda_min, double lambda_max, int n_lambda, int powell_path, double *beta_out, double *coef0_out,
                  double *train_loss_out, double *ic_out, double *lambda_out, double *rec_out, int max_rec)
{
    Eigen::MatrixXd X = Pointer2MatrixXd(x, n, p);
    Eigen::VectorXd Y = Pointer2VectorXd(y, n);
    Eigen::VectorXd W = Pointer2VectorXd(weight, n);
    Eigen::VectorXi g_index = Eigen::VectorXi::LinSpaced(p, 0, p - 1);
    srand(123);  // bess.cpp:53
    Data data(X, Y, data_type, W, is_normal, g_index);
    FitLog flog;
    Algorithm *alg;
    if (model_type == 1) { data.add_weight(); alg = new Traced<GroupPdasLm>(data, algorithm_type, max_iter, &flog); }
    else if (model_type == 2) alg = new Traced<GroupPdasLogistic>(data, algorithm_type, max_iter, &flog);
    else if (model_type == 3) alg = new Traced<GroupPdasPoisson>(data, algorithm_type, max_iter, &flog);
    else alg = new Traced<GroupPdasCox>(data, algorithm_type, max_iter, &flog);
    alg->set_warm_start(is_warm_start);
    alg->always_select = Eigen::VectorXi(0);
    alg->tao = 1.1;
    Metric *metric = make_metric(model_type, ic_type, is_cv, K);
    if (is_cv) {
        metric->set_cv_train_test_mask(data.get_n());
        metric->set_cv_initial_model_param(K, data.get_p());
        if (model_type == 1) metric->cal_cv_group_XTX(data);
    }
    List res = pgs_path(data, alg, metric, s_min, s_max, log(std::max(lambda_min, 1e-5)), log(std::max(lambda_max, 1e-5)),
                        powell_path, n_lambda);  // bess.cpp:171-174
    Eigen::VectorXd beta;
    res.get_value_by_name("beta", beta);
    res.get_value_by_name("coef0", *coef0_out);
    res.get_value_by_name("train_loss", *train_loss_out);
    res.get_value_by_name("ic", *ic_out);
    res.get_value_by_name("lambda", *lambda_out);
    for (int j = 0; j < p; j++) beta_out[j] = beta(j);
    const int nfits = (int)(flog.rec.size() / 5);
    for (int i = 0; i < std::min(nfits, max_rec) * 5; i++) rec_out[i] = flog.rec[(size_t)i];
    delete alg;
    delete metric;
    return nfits;
}

// max_k itself (utilities.cpp:179-188): the reference's resolution of boundary ties (std::nth_element over an index array)
void ref_max_k(const double *v, int n, int k, int *out)
{
    Eigen::VectorXd vec = Eigen::Map<const Eigen::VectorXd>(v, n);
    Eigen::VectorXi res;
    max_k(vec, k, res);
    for (int i = 0; i < k; i++) out[i] = res(i);
}

}  // extern "C"
