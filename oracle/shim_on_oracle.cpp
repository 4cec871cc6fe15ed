// shim_on_oracle.cpp -- TEST-ONLY artifact: the product's host shim
// (swarmkit_b200/csrc/scheduler_host.cpp) compiled against the CPU oracle's
// implementation of the C ABI (ope_* in flat_oracle.cpp) instead of the CUDA
// engine.  It lets the CPU test-suite (no GPU) check the shim's encoder and
// cross-check the encoded oracle against the object-level oracle.  It is built
// into oracle/_build/ and is never shipped, imported or loaded by the product.
#define pe_abi_version ope_abi_version
#define pe_create ope_create
#define pe_destroy ope_destroy
#define pe_last_error ope_last_error
#define pe_node_upsert ope_node_upsert
#define pe_node_remove ope_node_remove
#define pe_set_node_count ope_set_node_count
#define pe_node_task_delta ope_node_task_delta
#define pe_schedule ope_schedule
#define pe_tick_upload ope_tick_upload
#define pe_tick_run ope_tick_run
#define pe_tick_download ope_tick_download
#define pe_fit ope_fit
#define pe_snapshot ope_snapshot
#define pe_snapshot_service ope_snapshot_service
#define pe_snapshot_generic ope_snapshot_generic
#define pe_snapshot_ports ope_snapshot_ports
#define pe_get_stats ope_get_stats
#define pe_stats_reset ope_stats_reset
#define pe_fold_value ope_fold_value
#define pe_nccl_unique_id ope_nccl_unique_id
#define pe_pref_leaves ope_pref_leaves
#define pe_match_matrix ope_match_matrix
#define ss_create sso_create
#define ss_destroy sso_destroy
#define ss_apply sso_apply
#define ss_free sso_free
#define ss_last_error sso_last_error
#include "../swarmkit_b200/csrc/scheduler_host.cpp"
