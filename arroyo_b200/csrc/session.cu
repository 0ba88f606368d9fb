// Session window aggregate -- implemented in a later milestone of this round.
#include "op.h"
namespace ab {
OpBase* make_session_op(const ArroyoB200OpConfig&) {
  throw Error(ARROYO_B200_UNSUPPORTED, "SessionWindowAggregate is not built yet: use the stock operator");
}
}  // namespace ab
