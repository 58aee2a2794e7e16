// hb_policy.cu -- R2D2 policy forward on the device (placeholder until the kernels land).
#include "hb_engine.h"
extern "C" {
int hb_policy_create(hb_engine* e) { e->policy = nullptr; return 0; }
void hb_policy_destroy(hb_engine* e) { (void)e; }
}
