// facsimile of src/TiledArray/math/gemm_helper.h:41-278
#pragma once
#include <cstdint>
namespace blas { enum class Op : char { NoTrans = 'N', Trans = 'T', ConjTrans = 'C' }; typedef std::int64_t integer; }
namespace TiledArray { namespace math {
class GemmHelper {
 public:
  GemmHelper(blas::Op left_op, blas::Op right_op, unsigned int result_rank, unsigned int left_rank, unsigned int right_rank);  // :62-98
  unsigned int num_contract_ranks() const;  // :115
  template <typename R, typename Left, typename Right> R make_result_range(const Left& left, const Right& right) const;  // :166-192
  template <typename Left, typename Right> bool left_right_congruent(const Left& left, const Right& right) const;        // :239
  template <typename Left, typename Right>
  void compute_matrix_sizes(blas::integer& m, blas::integer& n, blas::integer& k, const Left& left, const Right& right) const;  // :255-274
  blas::Op left_op() const;   // :276
  blas::Op right_op() const;  // :277
};
}}  // namespace TiledArray::math
