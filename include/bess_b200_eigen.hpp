/* bess_b200_eigen.hpp -- the reference's C++ entry point, `bessCpp`, as a header-only adaptor over the C ABI.
 *
 * Replaces /root/reference/src/bess.h:20-33 / bess.cpp:37-214: the same thirty arguments in the same order, all by value, the
 * same result container (a name -> value bag with the keys "beta", "coef0", "train_loss", "ic", "lambda" and, after
 * screening, "screening_A", exactly what src/List.h:10-36 holds when R_BUILD is undefined).  An Rcpp build (R/src/
 * RcppExports.cpp:11-48) binds the same function; with R_BUILD defined there the container is Rcpp::List instead -- the
 * stub for that is in INTEGRATION.md.
 *
 * Header-only on purpose: it needs Eigen (any 3.3+; the reference vendors 3.3.4 under python/include), the library itself
 * does not.  Nothing in here computes: x is repacked row-major (the layout pywrap_bess receives, utilities.cpp:13-25) and
 * handed to bess_b200_fit; there is no CPU fallback -- without an sm_100a device the call throws std::runtime_error.
 *
 *     #include <bess_b200_eigen.hpp>
 *     bess_b200::List r = bess_b200::bessCpp(x, y, 1, w, true, 1, 1, 20, 2, 1, true, 3, false, 5, state, seq, lam,
 *                                            1, 20, 10, 10.0, 0.0, 0.0, 1, false, 1, 1, g_index, always, 1.1);
 *     Eigen::VectorXd beta; r.get_value_by_name("beta", beta);
 */
#ifndef BESS_B200_EIGEN_HPP
#define BESS_B200_EIGEN_HPP

#include <Eigen/Eigen>
#include <stdexcept>
#include <string>
#include <vector>

#include "bess_b200.h"

namespace bess_b200 {

/* src/List.h:10-36: add(name, value) / get_value_by_name(name, value&) for int, double, MatrixXd, VectorXd, VectorXi.
 * Like the reference's, a name that was added twice yields its LAST value (bess.cpp:199 re-adds "beta" after un-screening)
 * and a name that is absent leaves `value` untouched. */
class List {
public:
    void add(const std::string &name, int value) { put(ints_, name, value); }
    void add(const std::string &name, double value) { put(doubles_, name, value); }
    void add(const std::string &name, const Eigen::MatrixXd &value) { put(mats_, name, value); }
    void add(const std::string &name, const Eigen::VectorXd &value) { put(vecs_, name, value); }
    void add(const std::string &name, const Eigen::VectorXi &value) { put(ivecs_, name, value); }
    void get_value_by_name(const std::string &name, int &value) const { get(ints_, name, value); }
    void get_value_by_name(const std::string &name, double &value) const { get(doubles_, name, value); }
    void get_value_by_name(const std::string &name, Eigen::MatrixXd &value) const { get(mats_, name, value); }
    void get_value_by_name(const std::string &name, Eigen::VectorXd &value) const { get(vecs_, name, value); }
    void get_value_by_name(const std::string &name, Eigen::VectorXi &value) const { get(ivecs_, name, value); }
    bool has(const std::string &name) const
    {
        return find(ints_, name) || find(doubles_, name) || find(mats_, name) || find(vecs_, name) || find(ivecs_, name);
    }

private:
    template <class T>
    using Bag = std::vector<std::pair<std::string, T>>;
    template <class T>
    static void put(Bag<T> &bag, const std::string &name, const T &v)
    {
        for (auto &e : bag)
            if (e.first == name) {
                e.second = v;
                return;
            }
        bag.emplace_back(name, v);
    }
    template <class T>
    static bool find(const Bag<T> &bag, const std::string &name)
    {
        for (const auto &e : bag)
            if (e.first == name) return true;
        return false;
    }
    template <class T>
    static void get(const Bag<T> &bag, const std::string &name, T &v)
    {
        for (const auto &e : bag)
            if (e.first == name) v = e.second;
    }
    Bag<int> ints_;
    Bag<double> doubles_;
    Bag<Eigen::MatrixXd> mats_;
    Bag<Eigen::VectorXd> vecs_;
    Bag<Eigen::VectorXi> ivecs_;
};

/* bess.h:20-33, argument for argument.  Extra keys next to the reference's: "s" (sparsity level of the returned model) and
 * the per-evaluation trace the R build returns (path.cpp:116-123): "s_all", "ic_all", "train_loss_all", "coef0_all" and,
 * for the sequential path, "beta_all" (p x evaluations, de-normalised). */
inline List bessCpp(Eigen::MatrixXd x, Eigen::VectorXd y, int data_type, Eigen::VectorXd weight, bool is_normal, int algorithm_type,
                    int model_type, int max_iter, int exchange_num, int path_type, bool is_warm_start, int ic_type, bool is_cv,
                    int K, Eigen::VectorXd state, Eigen::VectorXi sequence, Eigen::VectorXd lambda_seq, int s_min, int s_max,
                    int K_max, double epsilon, double lambda_min, double lambda_max, int nlambda, bool is_screening,
                    int screening_size, int powell_path, Eigen::VectorXi g_index, Eigen::VectorXi always_select, double tao,
                    const bess_b200_ext *ext_in = nullptr)
{
    const int n = (int)x.rows(), p = (int)x.cols();
    Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> xr = x;  // utilities.cpp:13-25 in reverse
    Eigen::VectorXd beta = Eigen::VectorXd::Zero(p);
    double coef0 = 0.0, train_loss = 0.0, ic = 0.0, lambda = 0.0;
    int chosen_s = 0;
    Eigen::VectorXi screening_A = Eigen::VectorXi::Zero(is_screening ? screening_size : 1);
    bess_b200_ext ext = ext_in ? *ext_in : bess_b200_ext();
    if (!ext_in) ext.device = -1;
    ext.screening_A_out = screening_A.data();
    ext.chosen_s_out = &chosen_s;
    ext.chosen_lambda_out = &lambda;
    ext.beta_out_zeroed = 1;
    const int rc = bess_b200_fit(xr.data(), n, p, y.data(), (int)y.size(), data_type, weight.data(), (int)weight.size(), is_normal,
                                 algorithm_type, model_type, max_iter, exchange_num, path_type, is_warm_start, ic_type, is_cv, K,
                                 g_index.size() ? g_index.data() : nullptr, (int)g_index.size(), state.data(), (int)state.size(),
                                 sequence.data(), (int)sequence.size(), lambda_seq.data(), (int)lambda_seq.size(), s_min, s_max,
                                 K_max, epsilon, lambda_min, lambda_max, nlambda, is_screening, screening_size, powell_path,
                                 always_select.size() ? always_select.data() : nullptr, (int)always_select.size(), tao,
                                 beta.data(), p, &coef0, &train_loss, &ic, &ext);
    if (rc != 0) throw std::runtime_error(std::string("bess_b200: ") + bess_b200_last_error());
    List result;
    result.add("beta", beta);
    result.add("coef0", coef0);
    result.add("train_loss", train_loss);
    result.add("ic", ic);
    result.add("lambda", lambda);
    if (is_screening) result.add("screening_A", screening_A);
    result.add("s", chosen_s);
    const int len = bess_b200_trace(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
    if (len > 0) {
        Eigen::VectorXi s_all(len), l_all(len);
        Eigen::VectorXd c_all(len), t_all(len), i_all(len);
        const bool full = path_type == 1 || algorithm_type == 3 || algorithm_type == 5;
        Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> b_all;
        if (full) b_all.resize(len, p);
        bess_b200_trace(s_all.data(), l_all.data(), full ? c_all.data() : nullptr, t_all.data(), i_all.data(),
                        full ? b_all.data() : nullptr, p);
        result.add("s_all", s_all);
        result.add("ic_all", i_all);
        result.add("train_loss_all", t_all);
        if (full) {
            result.add("coef0_all", c_all);
            result.add("beta_all", Eigen::MatrixXd(b_all.transpose()));
        }
    }
    return result;
}

}  // namespace bess_b200

#endif /* BESS_B200_EIGEN_HPP */
