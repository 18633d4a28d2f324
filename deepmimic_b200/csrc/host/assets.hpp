// Host-side asset layer: loads the reference's on-disk formats into plain structs.
//   character  .txt : Skeleton.Joints / BodyDefs        (R/DeepMimicCore/anim/KinTree.cpp:204-258,439-487,852-916)
//   controller .txt : PDControllers + ct flags           (R/DeepMimicCore/sim/PDController.cpp:7-97, sim/CtController.cpp:161-172)
//   motion     .txt : Loop / Frames                      (R/DeepMimicCore/anim/Motion.cpp:104-141,303-430)
//   arg files       : scene keys                         
// Only parsing and the load-time post-processing the reference does once (quaternion
// normalisation, cumulative frame times, x/z recentring).  No simulation arithmetic here.
#pragma once
#include <algorithm>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

#include "arg_parser.hpp"
#include "hmath.hpp"
#include "json.hpp"

namespace dmh {

enum JointType { kRevolute = 0, kPlanar = 1, kPrismatic = 2, kFixed = 3, kSpherical = 4, kNone = 5 };
enum ShapeType { kShapeNull = 0, kShapeBox = 1, kShapeCapsule = 2, kShapeSphere = 3, kShapeCylinder = 4, kShapePlane = 5 };

struct JointDef {
    std::string name;
    int type = kRevolute;
    int parent = -1;
    V3 attach_pt;     // in parent joint frame (root forced to 0, KinTree.cpp:1017-1019)
    V3 attach_theta;  // Euler, R = Rz Ry Rx
    double lim_low[3] = {1, 1, 1};
    double lim_high[3] = {0, 0, 0};
    double torque_lim = std::numeric_limits<double>::infinity();
    double force_lim = std::numeric_limits<double>::infinity();
    bool is_end_eff = false;
    double diff_weight = 1;
    int param_offset = 0;  // into pose / vel vectors
    int param_size = 0;
};

struct BodyDef {
    std::string name;
    int shape = kShapeNull;
    double mass = 0;
    int col_group = -1;
    bool fall_contact = false;
    V3 attach_pt;     // body COM in its joint frame
    V3 attach_theta;  // body frame orientation in its joint frame
    double param[3] = {0, 0, 0};  // box: full extents; capsule: diameter, cyl-height; sphere: diameter
};

struct PDParams {
    double kp = 0, kd = 0;
    double target_theta[7] = {0, 0, 0, 0, 0, 0, 0};
    bool use_world_coord = false;
};

struct CharModel {
    std::vector<JointDef> joints;
    std::vector<BodyDef> bodies;
    int pose_dim = 0;  // == vel dim (43 humanoid3d, 83 dog3d)
    int num_joints() const { return static_cast<int>(joints.size()); }
    double total_mass() const {
        double m = 0;
        for (const auto& b : bodies)
            if (b.shape != kShapeNull) m += b.mass;
        return m;
    }
};

struct CtrlParams {
    std::vector<PDParams> pd;  // one per joint
    double query_rate = 30.0;  // key "QueryRate"; the shipped files say "UpdateRate", which is ignored
    double cycle_period = 1.0;
    bool enable_phase_input = false;
    bool record_world_root_pos = false;
    bool record_world_root_rot = false;
};

struct MotionClip {
    bool loop = false;
    int num_frames = 0;
    int frame_size = 0;
    std::vector<double> frame_times;  // cumulative start time of each frame
    std::vector<double> frames;       // num_frames x frame_size, quats normalised, root x/z recentred
    double duration() const { return frame_times.empty() ? 0 : frame_times.back(); }
    const double* frame(int f) const { return &frames[static_cast<size_t>(f) * frame_size]; }
};

struct SceneConfig {
    std::string scene = "imitate";
    int num_update_substeps = 1;
    int num_sim_substeps = 1;
    double world_scale = 1;
    V3 gravity = V3(0, -9.8, 0);
    std::string character_file, ctrl_file, motion_file, terrain_file, char_ctrl, kin_ctrl;
    std::vector<int> fall_contact_bodies;
    bool enable_char_contact_fall = true;
    bool enable_fall_end = true;
    bool sync_char_root_pos = true;
    bool sync_char_root_rot = false;
    bool enable_rand_rot_reset = false;
    bool enable_root_rot_fail = false;
    bool enable_amp_obs_local_root = false;
    double time_lim_min = std::numeric_limits<double>::infinity();
    double time_lim_max = std::numeric_limits<double>::infinity();
    double time_end_lim_min = std::numeric_limits<double>::infinity();
    double time_end_lim_max = std::numeric_limits<double>::infinity();
    int anneal_samples = -1;
    // AMP task scenes (cSceneTargetAMP::ParseArgs SceneTargetAMP.cpp:107-120, cSceneHeadingAMP::ParseArgs SceneHeadingAMP.cpp:71-88);
    // defaults are the constructors' (SceneTargetAMP.cpp:89-101, SceneHeadingAMP.cpp:50-65)
    double rand_target_time_min = 1, rand_target_time_max = 5;
    double max_target_dist = 3, target_succ_dist = 0.5, tar_fail_dist = std::numeric_limits<double>::infinity();
    double tar_speed = 1, pos_reward_scale = 1;
    bool enable_min_tar_vel = false;
    double max_heading_turn_rate = 0.15, sharp_turn_prob = 0.025, speed_change_prob = 0.1;
    double tar_speed_min = 1, tar_speed_max = 1, vel_reward_scale = 1;
    // cSceneHeadingAMPGetup::ParseArgs (SceneHeadingAMPGetup.cpp:76-85; constructor defaults :60-70)
    std::vector<int> getup_motion_ids;
    double getup_height_root = 0.5, getup_height_head = 0.5, recover_episode_prob = 0.0;
    int head_id = 0;
    // cSceneStrikeAMP::ParseArgs (SceneStrikeAMP.cpp:212-229; constructor defaults :192-206)
    V3 target_min = V3(-0.5, 1.2, 0.6), target_max = V3(0.5, 1.4, 1.1);
    double target_radius = 0.2, target_hit_reset_time = 2.0, tar_reward_scale = 2.0, hit_tar_speed = 1.5, init_hit_prob = 0.0, tar_far_prob = 0.4, tar_near_dist = 1.4;
    std::vector<int> strike_bodies, fail_tar_contact_bodies;
    bool is_task_scene() const { return scene == "target_amp" || scene == "heading_amp" || scene == "heading_amp_getup" || scene == "strike_amp"; }
    bool is_heading_scene() const { return scene == "heading_amp" || scene == "heading_amp_getup"; }
};

inline int joint_param_size(int type, bool is_root) {
    if (is_root) return 7;
    switch (type) {
        case kRevolute: return 1;
        case kPrismatic: return 1;
        case kPlanar: return 3;
        case kFixed: return 0;
        case kSpherical: return 4;
        default: return 0;
    }
}

inline int parse_joint_type(const std::string& s) {
    static const char* names[] = {"revolute", "planar", "prismatic", "fixed", "spherical", "none"};
    for (int i = 0; i < 6; ++i)
        if (s == names[i]) return i;
    throw std::runtime_error("Unsupported joint type: " + s);
}

inline int parse_shape(const std::string& s) {
    static const char* names[] = {"null", "box", "capsule", "sphere", "cylinder", "plane"};
    for (int i = 0; i < 6; ++i)
        if (s == names[i]) return i;
    throw std::runtime_error("Unsupported body shape " + s);
}

inline std::string resolve_path(const std::string& root, const std::string& p) {
    if (p.empty() || p[0] == '/' || root.empty()) return p;
    return root + "/" + p;
}

inline CharModel load_character(const std::string& path) {
    Json root = Json::parseFile(path);
    const Json& skel = root["Skeleton"];
    if (skel.isNull() || skel["Joints"].isNull()) throw std::runtime_error("Failed to parse character from file " + path);
    CharModel cm;
    const Json& joints = skel["Joints"];
    for (size_t j = 0; j < joints.size(); ++j) {
        const Json& jj = joints[j];
        JointDef d;
        d.name = jj["Name"].asString();
        if (jj["Type"].isNull()) std::printf("No joint type specified\n");
        else d.type = parse_joint_type(jj["Type"].asString());
        d.parent = jj["Parent"].isNull() ? -1 : jj["Parent"].asInt();
        d.attach_pt = V3(jj.get("AttachX", 0), jj.get("AttachY", 0), jj.get("AttachZ", 0));
        d.attach_theta = V3(jj.get("AttachThetaX", 0), jj.get("AttachThetaY", 0), jj.get("AttachThetaZ", 0));
        static const char* lo[] = {"LimLow0", "LimLow1", "LimLow2"};
        static const char* hi[] = {"LimHigh0", "LimHigh1", "LimHigh2"};
        for (int k = 0; k < 3; ++k) {
            d.lim_low[k] = jj.get(lo[k], 1);
            d.lim_high[k] = jj.get(hi[k], 0);
        }
        d.torque_lim = jj.get("TorqueLim", std::numeric_limits<double>::infinity());
        d.force_lim = jj.get("ForceLim", std::numeric_limits<double>::infinity());
        d.is_end_eff = jj.get("IsEndEffector", 0) != 0;
        d.diff_weight = jj.get("DiffWeight", 1);
        if (d.parent >= static_cast<int>(j))
            throw std::runtime_error("Parent id must be < child id in " + path);
        cm.joints.push_back(d);
    }
    int offset = 0;
    for (size_t j = 0; j < cm.joints.size(); ++j) {
        JointDef& d = cm.joints[j];
        d.param_size = joint_param_size(d.type, d.parent < 0);
        d.param_offset = offset;
        offset += d.param_size;
    }
    cm.pose_dim = offset;
    if (!cm.joints.empty()) cm.joints[0].attach_pt = V3(0, 0, 0);

    const Json& bodies = root["BodyDefs"];
    if (bodies.isNull()) throw std::runtime_error("Failed to load body definition from " + path);
    for (size_t b = 0; b < bodies.size(); ++b) {
        const Json& bj = bodies[b];
        BodyDef d;
        d.name = bj["Name"].asString();
        d.shape = parse_shape(bj["Shape"].asString("null"));
        d.mass = bj.get("Mass", 0);
        d.col_group = static_cast<int>(bj.get("ColGroup", -1));
        d.fall_contact = bj.get("EnableFallContact", 0) != 0;
        d.attach_pt = V3(bj.get("AttachX", 0), bj.get("AttachY", 0), bj.get("AttachZ", 0));
        d.attach_theta = V3(bj.get("AttachThetaX", 0), bj.get("AttachThetaY", 0), bj.get("AttachThetaZ", 0));
        d.param[0] = bj.get("Param0", 0);
        d.param[1] = bj.get("Param1", 0);
        d.param[2] = bj.get("Param2", 0);
        cm.bodies.push_back(d);
    }
    if (cm.bodies.size() != cm.joints.size()) throw std::runtime_error("joint / body count mismatch in " + path);
    return cm;
}

inline CtrlParams load_controller(const std::string& path, const CharModel& cm) {
    Json root = Json::parseFile(path);
    CtrlParams cp;
    cp.query_rate = root.get("QueryRate", cp.query_rate);
    cp.cycle_period = root.get("CyclePeriod", cp.cycle_period);
    cp.enable_phase_input = root.getBool("EnablePhaseInput", false);
    cp.record_world_root_pos = root.getBool("RecordWorldRootPos", false);
    cp.record_world_root_rot = root.getBool("RecordWorldRootRot", false);
    const Json& pds = root["PDControllers"];
    if (pds.isNull()) throw std::runtime_error("Failed to initialize Ct-PD controller from " + path);
    if (static_cast<int>(pds.size()) != cm.num_joints()) throw std::runtime_error("PDControllers count mismatch in " + path);
    static const char* tk[] = {"TargetTheta0", "TargetTheta1", "TargetTheta2", "TargetTheta3",
                               "TargetTheta4", "TargetTheta5", "TargetTheta6"};
    for (size_t i = 0; i < pds.size(); ++i) {
        const Json& pj = pds[i];
        PDParams p;
        p.kp = pj.get("Kp", 0);
        p.kd = pj.get("Kd", 0);
        for (int k = 0; k < 7; ++k) p.target_theta[k] = pj.get(tk[k], 0);
        p.use_world_coord = pj.get("UseWorldCoord", 0) != 0;
        cp.pd.push_back(p);
    }
    return cp;
}

// Loads a clip and applies the reference's load-time post-processing:
//  cMotion::PostProcessFrames (Motion.cpp:403-430): durations -> cumulative times, root x/z shifted so
//  frame 0 sits at the origin, cKinTree::PostProcessPose (KinTree.cpp:1318-1334) normalises quaternions;
//  cKinController::PostProcessMotion (KinController.cpp:131-147) re-centres x/z again (a no-op after the first).
inline MotionClip load_motion(const std::string& path, const CharModel& cm) {
    Json root = Json::parseFile(path);
    MotionClip mc;
    if (!root["Loop"].isNull()) {
        std::string s = root["Loop"].asString();
        if (s == "wrap") mc.loop = true;
        else if (s == "none") mc.loop = false;
        else throw std::runtime_error("Unsupported loop mode: " + s);
    }
    const Json& frames = root["Frames"];
    if (!frames.isArray() || frames.size() == 0) throw std::runtime_error("Failed to load motion from file " + path);
    mc.num_frames = static_cast<int>(frames.size());
    mc.frame_size = static_cast<int>(frames[0].size()) - 1;
    if (mc.frame_size != cm.pose_dim) throw std::runtime_error("DOF mismatch between character and motion " + path);
    mc.frame_times.resize(mc.num_frames);
    mc.frames.resize(static_cast<size_t>(mc.num_frames) * mc.frame_size);
    std::vector<double> dur(mc.num_frames);
    for (int f = 0; f < mc.num_frames; ++f) {
        const Json& fj = frames[f];
        if (static_cast<int>(fj.size()) != mc.frame_size + 1) throw std::runtime_error("ragged frame in " + path);
        dur[f] = fj[0].asDouble();
        for (int k = 0; k < mc.frame_size; ++k) mc.frames[static_cast<size_t>(f) * mc.frame_size + k] = fj[k + 1].asDouble();
    }
    double t = 0;
    double off_x = mc.frames[0], off_z = mc.frames[2];
    for (int f = 0; f < mc.num_frames; ++f) {
        mc.frame_times[f] = t;
        t += dur[f];
        double* fr = &mc.frames[static_cast<size_t>(f) * mc.frame_size];
        fr[0] -= off_x;
        fr[2] -= off_z;
        auto normalize4 = [](double* q) {
            double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            for (int k = 0; k < 4; ++k) q[k] /= n;
        };
        normalize4(fr + 3);
        for (int j = 1; j < cm.num_joints(); ++j)
            if (cm.joints[j].type == kSpherical) normalize4(fr + cm.joints[j].param_offset);
    }
    return mc;
}

inline SceneConfig parse_scene_config(const ArgParser& ap) {
    SceneConfig sc;
    ap.ParseString("scene", sc.scene);
    ap.ParseInt("num_update_substeps", sc.num_update_substeps);
    ap.ParseInt("num_sim_substeps", sc.num_sim_substeps);
    ap.ParseDouble("world_scale", sc.world_scale);
    std::vector<double> g;
    if (ap.ParseDoubles("gravity", g))
        for (size_t i = 0; i < std::min<size_t>(3, g.size()); ++i) sc.gravity[static_cast<int>(i)] = g[i];
    std::vector<std::string> v;
    if (ap.ParseStrings("character_files", v) && !v.empty()) sc.character_file = v[0];
    if (ap.ParseStrings("char_ctrl_files", v) && !v.empty()) sc.ctrl_file = v[0];
    if (ap.ParseStrings("char_ctrls", v) && !v.empty()) sc.char_ctrl = v[0];
    ap.ParseString("kin_ctrl", sc.kin_ctrl);
    ap.ParseString("motion_file", sc.motion_file);
    ap.ParseString("terrain_file", sc.terrain_file);
    ap.ParseInts("fall_contact_bodies", sc.fall_contact_bodies);
    ap.ParseBool("enable_char_contact_fall", sc.enable_char_contact_fall);
    ap.ParseBool("enable_fall_end", sc.enable_fall_end);
    ap.ParseBool("sync_char_root_pos", sc.sync_char_root_pos);
    ap.ParseBool("sync_char_root_rot", sc.sync_char_root_rot);
    ap.ParseBool("enable_rand_rot_reset", sc.enable_rand_rot_reset);
    ap.ParseBool("enable_root_rot_fail", sc.enable_root_rot_fail);
    ap.ParseBool("enable_amp_obs_local_root", sc.enable_amp_obs_local_root);
    ap.ParseDouble("time_lim_min", sc.time_lim_min);
    ap.ParseDouble("time_lim_max", sc.time_lim_max);
    sc.time_end_lim_min = sc.time_lim_min;
    sc.time_end_lim_max = sc.time_lim_max;
    ap.ParseDouble("time_end_lim_min", sc.time_end_lim_min);
    ap.ParseDouble("time_end_lim_max", sc.time_end_lim_max);
    ap.ParseInt("anneal_samples", sc.anneal_samples);
    if (sc.is_heading_scene()) { sc.rand_target_time_min = 0.2; sc.rand_target_time_max = 0.5; }
    ap.ParseDouble("rand_target_time_min", sc.rand_target_time_min);
    ap.ParseDouble("rand_target_time_max", sc.rand_target_time_max);
    ap.ParseDouble("max_target_dist", sc.max_target_dist);
    ap.ParseDouble("target_succ_dist", sc.target_succ_dist);
    ap.ParseDouble("tar_fail_dist", sc.tar_fail_dist);
    ap.ParseDouble("tar_speed", sc.tar_speed);
    ap.ParseBool("enable_min_tar_vel", sc.enable_min_tar_vel);
    ap.ParseDouble("pos_reward_scale", sc.pos_reward_scale);
    ap.ParseDouble("max_heading_turn_rate", sc.max_heading_turn_rate);
    ap.ParseDouble("sharp_turn_prob", sc.sharp_turn_prob);
    ap.ParseDouble("speed_change_prob", sc.speed_change_prob);
    sc.tar_speed_min = sc.tar_speed_max = sc.tar_speed;
    ap.ParseDouble("tar_speed_min", sc.tar_speed_min);
    ap.ParseDouble("tar_speed_max", sc.tar_speed_max);
    if (sc.is_heading_scene()) sc.tar_speed = std::min(std::max(sc.tar_speed, sc.tar_speed_min), sc.tar_speed_max);   // SceneHeadingAMP.cpp:85
    ap.ParseDouble("vel_reward_scale", sc.vel_reward_scale);
    {
        std::vector<double> v3;
        if (ap.ParseDoubles("target_min", v3) && v3.size() >= 3) sc.target_min = V3(v3[0], v3[1], v3[2]);
        if (ap.ParseDoubles("target_max", v3) && v3.size() >= 3) sc.target_max = V3(v3[0], v3[1], v3[2]);
    }
    ap.ParseDouble("target_radius", sc.target_radius);
    ap.ParseDouble("target_hit_reset_time", sc.target_hit_reset_time);
    ap.ParseDouble("tar_reward_scale", sc.tar_reward_scale);
    ap.ParseDouble("hit_tar_speed", sc.hit_tar_speed);
    ap.ParseDouble("init_hit_prob", sc.init_hit_prob);
    ap.ParseDouble("tar_far_prob", sc.tar_far_prob);
    ap.ParseDouble("tar_near_dist", sc.tar_near_dist);
    ap.ParseInts("strike_bodies", sc.strike_bodies);
    ap.ParseInts("fail_tar_contact_bodies", sc.fail_tar_contact_bodies);
    ap.ParseInts("getup_motion_ids", sc.getup_motion_ids);
    ap.ParseDouble("getup_height_root", sc.getup_height_root);
    ap.ParseDouble("getup_height_head", sc.getup_height_head);
    ap.ParseInt("head_id", sc.head_id);
    ap.ParseDouble("recover_episode_prob", sc.recover_episode_prob);
    return sc;
}

// Everything one scene needs, loaded relative to `asset_root` (the directory that holds data/ and args/).
struct SceneAssets {
    SceneConfig cfg;
    CharModel character;
    CtrlParams ctrl;
    MotionClip motion;                 // the clip of --kin_ctrl motion; clips[0] with --kin_ctrl clips
    // --kin_ctrl clips: the dataset of cClipsController (anim/ClipsController.cpp:114-214): clips, their weights and the sampling CDF
    std::vector<MotionClip> clips;
    std::vector<double> clip_weights, clip_cdf;
    std::vector<std::string> clip_files;
    // cClipsController::SelectNewMotion (ClipsController.cpp:226-236): upper_bound of a uniform draw in the CDF
    int select_clip(double u01) const {
        auto it = std::upper_bound(clip_cdf.begin(), clip_cdf.end(), u01);
        return std::min(static_cast<int>(it - clip_cdf.begin()), static_cast<int>(clip_cdf.size()) - 1);
    }
};

// cClipsController::LoadParams / LoadMotions / BuildClipsCDF (anim/ClipsController.cpp:114-214)
inline void load_clip_dataset(const std::string& path, const std::string& asset_root, SceneAssets& sa) {
    Json root = Json::parseFile(path);
    const Json& motions = root["Motions"];
    if (!motions.isArray() || motions.size() == 0) throw std::runtime_error("Failed to load clips controller parameters from file " + path);
    double sum = 0;
    for (size_t i = 0; i < motions.size(); ++i) {
        const Json& e = motions[i];
        const std::string file = e["File"].asString();
        sa.clip_files.push_back(file);
        sa.clips.push_back(load_motion(resolve_path(asset_root, file), sa.character));
        const double w = e.get("Weight", 1.0);
        sa.clip_weights.push_back(w);
        sum += w;
        sa.clip_cdf.push_back(sum);
    }
    for (auto& c : sa.clip_cdf) c /= sum;
}

inline SceneAssets load_scene_assets(const ArgParser& ap, const std::string& asset_root) {
    SceneAssets sa;
    sa.cfg = parse_scene_config(ap);
    if (sa.cfg.character_file.empty()) throw std::runtime_error("No valid character file specified.");
    sa.character = load_character(resolve_path(asset_root, sa.cfg.character_file));
    if (!sa.cfg.fall_contact_bodies.empty()) {  // cSceneSimChar::SetFallContacts (SceneSimChar.cpp:460-476)
        for (auto& b : sa.character.bodies) b.fall_contact = false;
        for (int b : sa.cfg.fall_contact_bodies)
            if (b >= 0 && b < sa.character.num_joints()) sa.character.bodies[b].fall_contact = true;
    }
    if (sa.cfg.ctrl_file.empty()) throw std::runtime_error("no --char_ctrl_files given");
    sa.ctrl = load_controller(resolve_path(asset_root, sa.cfg.ctrl_file), sa.character);
    if (sa.cfg.motion_file.empty()) throw std::runtime_error("no --motion_file given");
    if (sa.cfg.kin_ctrl == "clips") {   // cKinCtrlBuilder::BuildClipsController (anim/KinCtrlBuilder.cpp:66-73)
        load_clip_dataset(resolve_path(asset_root, sa.cfg.motion_file), asset_root, sa);
        sa.motion = sa.clips[0];
    } else {
        sa.motion = load_motion(resolve_path(asset_root, sa.cfg.motion_file), sa.character);
        sa.clips.push_back(sa.motion); sa.clip_weights.push_back(1.0); sa.clip_cdf.push_back(1.0); sa.clip_files.push_back(sa.cfg.motion_file);
    }
    return sa;
}

}  // namespace dmh
