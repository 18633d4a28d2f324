// Host shim over deepmimic_b200/csrc/kernels/dm_task.cuh for tests/test_task_scenes_cpu.py: the per-environment task-scene logic that
// dm_step_kernel<.., TASK> and dm_task_*_kernel run on the device, compiled here with g++ so it can be checked against the oracle on the CPU.
#include "../deepmimic_b200/csrc/kernels/dm_task.cuh"

using namespace dmk;

static TaskParams params_from(const double* p) {
    TaskParams P;
    P.timer_min = p[1]; P.timer_max = p[2]; P.max_target_dist = p[3]; P.target_succ_dist = p[4]; P.tar_fail_dist = p[5]; P.pos_reward_scale = p[6];
    P.max_heading_turn_rate = p[7]; P.sharp_turn_prob = p[8]; P.speed_change_prob = p[9]; P.tar_speed_min = p[10]; P.tar_speed_max = p[11]; P.vel_reward_scale = p[12];
    P.tar_speed = p[13]; P.enable_min_tar_vel = static_cast<int>(p[14]); P.pad_ = 0;
    return P;
}

extern "C" {
// p: the 16 doubles of dm_get_task_params (p[0] = kind); t: the environment's task block (kTaskDoubles)
void shim_reset(const double* p, double* t, unsigned long long seed, unsigned long long env, double root_x, double root_z) {
    TaskRng r{seed, env, t + kKCounter};
    task_reset(static_cast<int>(p[0]), params_from(p), t, r, root_x, root_z);
}
void shim_update(const double* p, double* t, unsigned long long seed, unsigned long long env, double dt, double root_x, double root_z) {
    TaskRng r{seed, env, t + kKCounter};
    task_update(static_cast<int>(p[0]), params_from(p), t, r, dt, root_x, root_z);
}
int shim_dist_fail(const double* p, const double* t, double root_x, double root_z) { return task_dist_fail(static_cast<int>(p[0]), params_from(p), t, root_x, root_z) ? 1 : 0; }
void shim_goal(const double* p, const double* t, double root_x, double root_z, double heading, double* out3) { task_goal(static_cast<int>(p[0]), t, root_x, root_z, heading, out3); }
double shim_reward(const double* p, const double* t, int fallen, double root_x, double root_z, double step_dur) {
    return task_reward(static_cast<int>(p[0]), params_from(p), t, fallen != 0, root_x, root_z, step_dur);
}
int shim_task_doubles() { return kTaskDoubles; }
// cSceneImitate::SyncKinCharNewCycle as the kVarRootRot instantiation of dm_step_kernel runs it
void shim_wrap_sync(const double* frame_times, const float* frames, int pose_dim, int num_frames, const float* cycle_delta, double dur, double kin_time,
                    double* origin, double* origin_rot, double sim_x, double sim_z, const double* sim_quat, int sync_pos, int sync_rot) {
    kin_wrap_sync(frame_times, frames, pose_dim, num_frames, cycle_delta, dur, kin_time, origin, origin_rot, sim_x, sim_z, sim_quat, sync_pos != 0, sync_rot != 0);
}
}
