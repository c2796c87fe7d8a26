/* tests/helpers/arith_host.cpp -- host build of the product's arithmetic
 * header for CPU-side fuzzing against oracle/_ref (test infrastructure). */
#include "../../saugns_b200/csrc/sau_arith.h"
extern "C" {
void arith_line_fill(uint32_t type, float *buf, uint32_t len, float v0, float vt,
		uint32_t pos, uint32_t time, const float *mulbuf, int tailrule) {
	sau::LineFill f = sau::line_fill_setup((int) type, v0, vt, pos, time);
	for (uint32_t i = 0; i < len; ++i) {
		bool tail = tailrule && (len & 1) && i == len - 1;
		float v = sau::line_fill_at(f, i, tail);
		buf[i] = mulbuf ? v * mulbuf[i] : v;
	}
}
float arith_line_val(uint32_t type, float x, float a, float b) {
	return sau::line_val((int) type, x, a, b, false);
}
void arith_line_map(uint32_t type, float *buf, uint32_t len, const float *e0, const float *e1, int tailrule) {
	for (uint32_t i = 0; i < len; ++i) {
		bool tail = tailrule && i >= (len & ~3u);   /* 4-wide body, scalar tail */
		buf[i] = sau::line_val((int) type, buf[i], e0[i], e1[i], tail);
	}
}
}
