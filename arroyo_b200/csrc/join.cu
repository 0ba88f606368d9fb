// Instant (windowed) join -- implemented in a later milestone of this round.
#include "op.h"
namespace ab {
OpBase* make_instant_join_op(const ArroyoB200OpConfig&) {
  throw Error(ARROYO_B200_UNSUPPORTED, "InstantJoin is not built yet: use the stock operator");
}
}  // namespace ab
