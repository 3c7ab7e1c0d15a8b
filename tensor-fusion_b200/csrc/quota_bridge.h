// quota_bridge.h -- moves tokens from the hypervisor's quota file into the
// device-resident bucket (DESIGN.md "limiter bridge").
#pragma once
#include <stdint.h>

#include "tfw_gate.h"

namespace tfw {

struct QuotaBridge;

tfw_status quota_bridge_start(tfw_gate* g, const char* shm_file, uint32_t device_index, QuotaBridge** out);
void quota_bridge_stop(QuotaBridge* b);
void quota_bridge_note_cost(QuotaBridge* b, double cost);  // largest single-launch cost seen
double quota_bridge_rate(QuotaBridge* b);
double quota_bridge_capacity(QuotaBridge* b);  // erl_token_capacity of the quota file as last read (0 = not read yet)
uint64_t quota_bridge_moved_milli(QuotaBridge* b);

// provided by gate.cu
int gate_device(tfw_gate* g);
double gate_mirror_tokens(tfw_gate* g);

}  // namespace tfw
