// hb_replay.cu -- device-resident prioritized episode replay (placeholder until the kernels land).
#include "hb_engine.h"
extern "C" {
int hb_replay_create(hb_engine* e) { e->replay = nullptr; return 0; }
void hb_replay_destroy(hb_engine* e) { (void)e; }
}
