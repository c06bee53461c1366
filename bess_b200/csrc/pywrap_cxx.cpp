// The reference exports pywrap_bess with C++ linkage (/root/reference/src/bess.h:35-51; mangled
// _Z11pywrap_bessPdiiS_iiS_ibiiiiibibiPiiS_iS0_iS_iiiidddibiiS0_idS_iS_iS_iS_iS_S_iS_iS_iS0_iS0_).  This translation unit
// exports the same symbol so a SWIG module generated from the reference's python/src/bess.i (or any C++ caller compiled
// against the reference's bess.h) links against libbess_b200.so unchanged.  It must live in its own file: the
// extern "C" twin in capi.cpp has the same name and parameter list.
void bess_b200_pywrap_impl(double *x, int x_row, int x_col, double *y, int y_len, int data_type, double *weight,
                           int weight_len, bool is_normal, int algorithm_type, int model_type, int max_iter,
                           int exchange_num, int path_type, bool is_warm_start, int ic_type, bool is_cv, int K,
                           int *gindex, int gindex_len, double *state, int state_len, int *sequence, int sequence_len,
                           double *lambda_sequence, int lambda_sequence_len, int s_min, int s_max, int K_max,
                           double epsilon, double lambda_min, double lambda_max, int n_lambda, bool is_screening,
                           int screening_size, int powell_path, int *always_select, int always_select_len, double tao,
                           double *beta_out, int beta_out_len, double *coef0_out, double *train_loss_out, double *ic_out);

void pywrap_bess(double *x, int x_row, int x_col, double *y, int y_len, int data_type, double *weight, int weight_len,
                 bool is_normal, int algorithm_type, int model_type, int max_iter, int exchange_num, int path_type,
                 bool is_warm_start, int ic_type, bool is_cv, int K, int *gindex, int gindex_len, double *state,
                 int state_len, int *sequence, int sequence_len, double *lambda_sequence, int lambda_sequence_len,
                 int s_min, int s_max, int K_max, double epsilon, double lambda_min, double lambda_max, int n_lambda,
                 bool is_screening, int screening_size, int powell_path, int *always_select, int always_select_len,
                 double tao, double *beta_out, int beta_out_len, double *coef0_out, int, double *train_loss_out, int,
                 double *ic_out, int, double *, double *, int, double *, int, double *, int, int *, int, int *)
{
    bess_b200_pywrap_impl(x, x_row, x_col, y, y_len, data_type, weight, weight_len, is_normal, algorithm_type, model_type,
                          max_iter, exchange_num, path_type, is_warm_start, ic_type, is_cv, K, gindex, gindex_len, state,
                          state_len, sequence, sequence_len, lambda_sequence, lambda_sequence_len, s_min, s_max, K_max,
                          epsilon, lambda_min, lambda_max, n_lambda, is_screening, screening_size, powell_path,
                          always_select, always_select_len, tao, beta_out, beta_out_len, coef0_out, train_loss_out, ic_out);
}
