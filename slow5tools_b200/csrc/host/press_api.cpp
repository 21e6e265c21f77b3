// press_api.cpp -- the stateful press objects of slow5_press.h:97-125 over the single-buffer C-ABI (include/slow5b200_press.h).
#include <cstdlib>
#include <cstring>

#include "../../../include/slow5b200_press.h"

namespace {
struct PressState {
    int footer_next;  // recorded for parity with the reference's flush flag; every call finishes its stream anyway
};
bool method_ok(int m) { return m >= S5B_COMPRESS_NONE && m <= S5B_COMPRESS_EX_ZD; }
}  // namespace

extern "C" {

struct __s5b_press *__s5b_press_init(int method) {
    if (!method_ok(method)) return nullptr;
    __s5b_press *p = static_cast<__s5b_press *>(calloc(1, sizeof *p));
    if (!p) return nullptr;
    p->method = method;
    p->stream = calloc(1, sizeof(PressState));
    if (!p->stream) {
        free(p);
        return nullptr;
    }
    return p;
}

void __s5b_press_free(struct __s5b_press *comp) {
    if (!comp) return;
    free(comp->stream);
    free(comp);
}

s5b_press_t *s5b_press_init(s5b_press_method_t method) {
    __s5b_press *rec = __s5b_press_init(method.record_method);
    __s5b_press *sig = __s5b_press_init(method.signal_method);
    s5b_press_t *c = static_cast<s5b_press_t *>(calloc(1, sizeof *c));
    if (!rec || !sig || !c) {
        __s5b_press_free(rec);
        __s5b_press_free(sig);
        free(c);
        return nullptr;
    }
    c->record_press = rec;
    c->signal_press = sig;
    return c;
}

void s5b_press_free(s5b_press_t *comp) {
    if (!comp) return;
    __s5b_press_free(comp->record_press);
    __s5b_press_free(comp->signal_press);
    free(comp);
}

void *s5b_ptr_compress(struct __s5b_press *comp, const void *ptr, size_t count, size_t *n) {
    // slow5_press.c:383-434: a NULL press object means "no compression"
    void *out = s5b_ptr_compress_solo(comp ? comp->method : S5B_COMPRESS_NONE, ptr, count, n);
    if (comp && comp->stream) static_cast<PressState *>(comp->stream)->footer_next = 0;
    return out;
}

void *s5b_ptr_depress(struct __s5b_press *comp, const void *ptr, size_t count, size_t *n) {
    if (!comp) {  // slow5_press.c:499-509: SLOW5_ERR_ARG
        if (n) *n = 0;
        return nullptr;
    }
    return s5b_ptr_depress_solo(comp->method, ptr, count, n);
}

void s5b_compress_footer_next(struct __s5b_press *comp) {
    if (comp && comp->stream) static_cast<PressState *>(comp->stream)->footer_next = 1;
}

}  // extern "C"
