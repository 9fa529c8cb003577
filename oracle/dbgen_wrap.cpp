// TEST INFRASTRUCTURE. C entry points over the reference's vendored TPC-H dbgen, compiled from the
// sources where they lie under /root/reference (oracle/build_ref.sh -> oracle/_ref/libtpchref.so).
// Follows the wrapper logic of velox/tpch/gen/TpchGen.cpp:402-490 (genTpchLineItem) and :492-534
// (genTpchPart), and the backend initialisation of velox/tpch/gen/DBGenIterator.cpp:30-60
// (load_dists / init_build_buffers, sd_order + sd_line seeding, row_start / mk_order / row_stop_h).
// Only the columns the hot path's queries read are exported.
#include <velox/tpch/gen/dbgen/include/dbgen/dbgen_gunk.hpp>
#include <velox/tpch/gen/dbgen/include/dbgen/dss.h>
#include <velox/tpch/gen/dbgen/include/dbgen/dsstypes.h>

#include <cstdint>
#include <cstring>
#include <mutex>

using namespace facebook::velox::tpch::dbgen;

namespace {
std::once_flag g_once;
void ensure_backend() {
  std::call_once(g_once, [] {
    DBGenContext ctx;
    load_dists(10 * 1024 * 1024, &ctx);  // text pool for comments (unused columns)
    init_build_buffers();
  });
}
// "yyyy-mm-dd" -> days since 1970-01-01 (DATE()->toDays in the reference wrapper)
int32_t to_days(const char* s) {
  int y = (s[0] - '0') * 1000 + (s[1] - '0') * 100 + (s[2] - '0') * 10 + (s[3] - '0');
  int m = (s[5] - '0') * 10 + (s[6] - '0');
  int d = (s[8] - '0') * 10 + (s[9] - '0');
  y -= m <= 2;
  const int era = (y >= 0 ? y : y - 399) / 400;
  const unsigned yoe = static_cast<unsigned>(y - era * 400);
  const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + static_cast<int>(doe) - 719468;
}
}  // namespace

extern "C" {

// Generates the lineitems of orders [order_offset, order_offset + n_orders) at `scale`.
// Output arrays must hold 7 * n_orders rows. Returns the number of lineitem rows written.
int64_t ref_gen_lineitem(double scale, int64_t order_offset, int64_t n_orders, int64_t* orderkey, int64_t* partkey, double* quantity,
                         double* extendedprice, double* discount, double* tax, char* returnflag, char* linestatus, int32_t* shipdate) {
  ensure_backend();
  DBGenContext ctx;
  ctx.scale_factor = scale < 1 && scale > 0 ? 1 : static_cast<long>(scale);
  sd_order(ORDER, order_offset, &ctx);
  sd_line(LINE, order_offset, &ctx);
  order_t order;
  int64_t n = 0;
  for (int64_t i = 0; i < n_orders; ++i) {
    row_start(ORDER, &ctx);
    mk_order(i + order_offset + 1, &order, &ctx, 0);
    row_stop_h(ORDER, &ctx);
    for (int64_t l = 0; l < order.lines; ++l) {
      const auto& line = order.l[l];
      orderkey[n] = line.okey;
      partkey[n] = line.partkey;
      quantity[n] = static_cast<double>(line.quantity);
      extendedprice[n] = static_cast<double>(line.eprice) * 0.01;
      discount[n] = static_cast<double>(line.discount) * 0.01;
      tax[n] = static_cast<double>(line.tax) * 0.01;
      returnflag[n] = line.rflag[0];
      linestatus[n] = line.lstatus[0];
      shipdate[n] = to_days(line.sdate);
      ++n;
    }
  }
  return n;
}

// Generates parts [offset, offset + n): partkey and p_type (type strings copied to a 26-byte slot each).
void ref_gen_part(double scale, int64_t offset, int64_t n, int64_t* partkey, char* type26) {
  ensure_backend();
  DBGenContext ctx;
  ctx.scale_factor = scale < 1 && scale > 0 ? 1 : static_cast<long>(scale);
  sd_part(PART, offset, &ctx);
  sd_psupp(PSUPP, offset, &ctx);
  part_t part;
  for (int64_t i = 0; i < n; ++i) {
    row_start(PART, &ctx);
    mk_part(i + offset + 1, &part, &ctx);
    row_stop_h(PART, &ctx);
    partkey[i] = part.partkey;
    std::memset(type26 + i * 26, 0, 26);
    std::strncpy(type26 + i * 26, part.type, 25);
  }
}

}  // extern "C"
