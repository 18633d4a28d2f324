// Host C++ mirror of the reference's SWIG-exported facade `cDeepMimicCore`
// (R/DeepMimicCore/DeepMimicCore.h:12-88, DeepMimicCore.i:1-35) over the C ABI of include/deepmimic_b200.h.
// Same method names, argument meaning and return shapes, so R/env/deepmimic_env.py runs unmodified:
//     from DeepMimicCore import DeepMimicCore ; core = DeepMimicCore.cDeepMimicCore(False)
// One facade object = one environment (the reference's shape) unless `--num_envs N` is given, in which case the
// per-agent getters address environment `agent_id` of the batch (an extension; the reference always passes 0).
// SWIG is not available in the build image, so the wrapper is pybind11 (deepmimic_b200/DeepMimicCore/).
// Draw / UI methods exist for surface compatibility and are no-ops: visualisation is outside the hot path.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/deepmimic_b200.h"

namespace py = pybind11;

class cDeepMimicCore {
public:
    explicit cDeepMimicCore(bool enable_draw) : mEnableDraw(false) {
        if (enable_draw) std::printf("[deepmimic_b200] enable_draw is ignored: rendering is outside the accelerated hot path\n");
    }
    virtual ~cDeepMimicCore() { if (mHandle) dm_destroy(mHandle); }

    virtual void SeedRand(int seed) { mSeed = static_cast<uint64_t>(seed); }
    virtual void ParseArgs(const std::vector<std::string>& args) {
        mArgs = args;
        for (size_t i = 0; i + 1 < args.size(); ++i) {
            if (args[i] == "--num_envs") mNumEnvs = std::atoi(args[i + 1].c_str());
            if (args[i] == "--asset_root") mAssetRoot = args[i + 1];
            if (args[i] == "--device") mDevice = std::atoi(args[i + 1].c_str());
        }
    }
    virtual void Init() {
        std::vector<const char*> argv;
        for (const auto& a : mArgs) argv.push_back(a.c_str());
        const char* env_root = std::getenv("DEEPMIMIC_ASSET_ROOT");
        std::string root = !mAssetRoot.empty() ? mAssetRoot : (env_root ? env_root : ".");
        mHandle = dm_create(root.c_str(), static_cast<int>(argv.size()), argv.data(), mNumEnvs, mDevice, mSeed, 0);
        if (!mHandle) Fatal(std::string("Init failed: ") + dm_last_error());   // the reference asserts (process abort) on bad arg files
        dm_get_dims(mHandle, &mDims);
        mState.assign(static_cast<size_t>(mNumEnvs) * mDims.state_size, 0.f);
        mReward.assign(mNumEnvs, 0.f);
        mFlags.assign(static_cast<size_t>(mNumEnvs) * 4, 0);
        mActions.assign(static_cast<size_t>(mNumEnvs) * mDims.action_size, 0.f);
        mActionDirty = false;
        Refresh();
    }
    virtual void Update(double timestep) {
        FlushActions();
        Check(dm_update(mHandle, timestep, 1));
        mTime += timestep;
        mFresh = false;
    }
    virtual void Reset() {
        // the reference resets its single scene; with a batch only finished episodes restart (all of them at N = 1)
        Refresh();
        bool any_done = false;
        for (int e = 0; e < mNumEnvs; ++e) any_done |= mFlags[4 * e + 1] != 0;
        Check(dm_reset(mHandle, (mNumEnvs == 1 || !any_done) ? 1 : 0, nullptr, nullptr, nullptr));
        mTime = 0;
        mFresh = false;
    }
    virtual double GetTime() const { return mTime; }
    virtual std::string GetName() const { char b[64] = "Imitate"; if (mHandle) dm_get_scene_name(mHandle, b, 64); return b; }
    virtual bool EnableDraw() const { return mEnableDraw; }

    virtual void Draw() {}
    virtual void Keyboard(int, int, int) {}
    virtual void MouseClick(int, int, int, int) {}
    virtual void MouseMove(int, int) {}
    virtual void Reshape(int, int) {}
    virtual void Shutdown() {}
    virtual bool IsDone() const { return false; }
    virtual void SetPlaybackSpeed(double speed) { mPlaybackSpeed = speed; }
    virtual void SetUpdatesPerSec(double ups) { mUpdatesPerSec = ups; }
    virtual int GetWinWidth() const { return 0; }
    virtual int GetWinHeight() const { return 0; }
    virtual int GetNumUpdateSubsteps() const { return mDims.num_update_substeps; }

    virtual bool IsRLScene() const { return true; }
    virtual int GetNumAgents() const { return 1; }
    virtual bool NeedNewAction(int agent_id) { Refresh(); return mFlags[4 * Env(agent_id) + 0] != 0; }
    virtual std::vector<double> RecordState(int agent_id) {
        Refresh();
        const float* s = &mState[static_cast<size_t>(Env(agent_id)) * mDims.state_size];
        return std::vector<double>(s, s + mDims.state_size);
    }
    virtual std::vector<double> RecordGoal(int agent_id) {
        if (mDims.goal_size == 0) return std::vector<double>();
        FlushActions();
        std::vector<float> buf(static_cast<size_t>(mNumEnvs) * mDims.goal_size);
        Check(dm_goal_host(mHandle, buf.data()));
        const float* p = &buf[static_cast<size_t>(Env(agent_id)) * mDims.goal_size];
        return std::vector<double>(p, p + mDims.goal_size);
    }
    virtual void SetAction(int agent_id, const std::vector<double>& action) {
        if (static_cast<int>(action.size()) != mDims.action_size) Fatal("SetAction: wrong action size");
        float* a = &mActions[static_cast<size_t>(Env(agent_id)) * mDims.action_size];
        for (int k = 0; k < mDims.action_size; ++k) a[k] = static_cast<float>(action[k]);
        mActionDirty = true;
    }
    virtual void LogVal(int, double) {}
    virtual int GetActionSpace(int) const { return 0; }   // eActionSpaceContinuous
    virtual int GetStateSize(int) const { return mDims.state_size; }
    virtual int GetGoalSize(int) const { return mDims.goal_size; }
    virtual int GetActionSize(int) const { return mDims.action_size; }
    virtual int GetNumActions(int) const { return 0; }
    virtual std::vector<double> BuildStateOffset(int) const { return Static(DM_STATE_OFFSET, mDims.state_size); }
    virtual std::vector<double> BuildStateScale(int) const { return Static(DM_STATE_SCALE, mDims.state_size); }
    int TaskKind() const {   // dm_task.cuh: 0 none, 1 target, 2 heading, 3 heading + get-up, 4 strike
        if (mDims.goal_size == 0) return 0;
        double p[48]; dm_get_task_params(mHandle, p, nullptr);
        return static_cast<int>(p[0]);
    }
    virtual std::vector<double> BuildGoalOffset(int) const {   // RLSceneSimChar.cpp:111-116; SceneHeadingAMPGetup.cpp:142-149
        std::vector<double> v(mDims.goal_size, 0.0);
        if (TaskKind() == 3) v[3] = -0.5;
        return v;
    }
    virtual std::vector<double> BuildGoalScale(int) const {
        std::vector<double> v(mDims.goal_size, 1.0);
        if (TaskKind() == 3) v[3] = 2.0;
        return v;
    }
    virtual std::vector<double> BuildActionOffset(int) const { return Static(DM_ACTION_OFFSET, mDims.action_size); }
    virtual std::vector<double> BuildActionScale(int) const { return Static(DM_ACTION_SCALE, mDims.action_size); }
    virtual std::vector<double> BuildActionBoundMin(int) const { return Static(DM_ACTION_BOUND_MIN, mDims.action_size); }
    virtual std::vector<double> BuildActionBoundMax(int) const { return Static(DM_ACTION_BOUND_MAX, mDims.action_size); }
    virtual std::vector<int> BuildStateNormGroups(int) const {
        std::vector<double> g = Static(DM_STATE_NORM_GROUPS, mDims.state_size);
        return std::vector<int>(g.begin(), g.end());
    }
    virtual std::vector<int> BuildGoalNormGroups(int) const {   // RLSceneSimChar.cpp:136-140; SceneHeadingAMPGetup.cpp:151-157; SceneStrikeAMP.cpp:401-405
        std::vector<int> v(mDims.goal_size, 0);
        const int kind = TaskKind();
        if (kind == 3) v[3] = -1;
        if (kind == 4) std::fill(v.begin(), v.end(), -1);
        return v;
    }
    virtual double CalcReward(int agent_id) { Refresh(); return mReward[Env(agent_id)]; }
    virtual double GetRewardMin(int) const { return 0; }
    virtual double GetRewardMax(int) const { return 1; }
    virtual double GetRewardFail(int) const { return 0; }
    virtual double GetRewardSucc(int) const { return 1; }
    virtual bool EnableAMPTaskReward() const { return mDims.goal_size > 0; }   // true in the task scenes (SceneTargetAMP.cpp:222-225)
    virtual int GetAMPObsSize() const { return mDims.amp_obs_size; }
    virtual std::vector<double> GetAMPObsOffset() const { return std::vector<double>(mDims.amp_obs_size, 0.0); }   // SceneImitateAMP.cpp:86-89
    virtual std::vector<double> GetAMPObsScale() const { return std::vector<double>(mDims.amp_obs_size, 1.0); }    // :91-94
    virtual std::vector<int> GetAMPObsNormGroup() const { return std::vector<int>(mDims.amp_obs_size, 0); }         // gNormGroupSingle, :96-99
    virtual std::vector<double> RecordAMPObsExpert(int agent_id) { return AmpObs(1, agent_id); }
    virtual std::vector<double> RecordAMPObsAgent(int agent_id) { return AmpObs(0, agent_id); }
    virtual bool IsEpisodeEnd() { Refresh(); for (int e = 0; e < mNumEnvs; ++e) if (mFlags[4 * e + 1]) return true; return false; }
    virtual bool CheckValidEpisode() { Refresh(); for (int e = 0; e < mNumEnvs; ++e) if (!mFlags[4 * e + 3]) return false; return true; }
    virtual int CheckTerminate(int agent_id) { Refresh(); return mFlags[4 * Env(agent_id) + 2]; }
    virtual void SetMode(int mode) { if (mHandle) dm_set_mode(mHandle, mode); }
    virtual void SetSampleCount(int count) { if (mHandle) dm_set_sample_count(mHandle, count); }
    // extension: batch size of this facade
    int GetNumEnvs() const { return mNumEnvs; }

private:
    [[noreturn]] void Fatal(const std::string& msg) const { std::fprintf(stderr, "[deepmimic_b200] %s\n", msg.c_str()); throw std::runtime_error(msg); }
    void Check(int rc) const { if (rc != 0) Fatal(dm_last_error()); }
    int Env(int agent_id) const { return (agent_id >= 0 && agent_id < mNumEnvs) ? agent_id : 0; }
    void FlushActions() {
        if (!mActionDirty) return;
        Check(dm_step_host(mHandle, mActions.data(), 0.0, 0, nullptr, nullptr, nullptr));
        mActionDirty = false;
    }
    void Refresh() {   // observation / reward / flags of the current state, fetched once per state
        if (mFresh) return;
        FlushActions();
        Check(dm_step_host(mHandle, nullptr, 0.0, 0, mState.data(), mReward.data(), mFlags.data()));
        mFresh = true;
    }
    std::vector<double> AmpObs(int expert, int agent_id) {
        FlushActions();
        std::vector<float> buf(static_cast<size_t>(mNumEnvs) * mDims.amp_obs_size);
        Check(dm_amp_obs_host(mHandle, expert, nullptr, buf.data()));
        const float* p = &buf[static_cast<size_t>(Env(agent_id)) * mDims.amp_obs_size];
        return std::vector<double>(p, p + mDims.amp_obs_size);
    }
    std::vector<double> Static(int kind, int n) const {
        std::vector<double> v(n);
        if (n > 0) dm_get_static(mHandle, kind, v.data());
        return v;
    }
    dm_handle* mHandle = nullptr;
    dm_dims mDims{};
    std::vector<std::string> mArgs;
    std::string mAssetRoot;
    int mNumEnvs = 1, mDevice = 0;
    uint64_t mSeed = 0;
    bool mEnableDraw, mFresh = false, mActionDirty = false;
    double mTime = 0, mPlaybackSpeed = 1, mUpdatesPerSec = 0;
    std::vector<float> mState, mReward, mActions;
    std::vector<int32_t> mFlags;
};

PYBIND11_MODULE(_DeepMimicCore, m) {
    m.doc() = "deepmimic_b200: cDeepMimicCore facade over the sm_100a batched step";
    py::class_<cDeepMimicCore>(m, "cDeepMimicCore")
        .def(py::init<bool>())
        .def("SeedRand", &cDeepMimicCore::SeedRand).def("ParseArgs", &cDeepMimicCore::ParseArgs).def("Init", &cDeepMimicCore::Init)
        .def("Update", &cDeepMimicCore::Update).def("Reset", &cDeepMimicCore::Reset).def("GetTime", &cDeepMimicCore::GetTime)
        .def("GetName", &cDeepMimicCore::GetName).def("EnableDraw", &cDeepMimicCore::EnableDraw).def("Draw", &cDeepMimicCore::Draw)
        .def("Keyboard", &cDeepMimicCore::Keyboard).def("MouseClick", &cDeepMimicCore::MouseClick).def("MouseMove", &cDeepMimicCore::MouseMove)
        .def("Reshape", &cDeepMimicCore::Reshape).def("Shutdown", &cDeepMimicCore::Shutdown).def("IsDone", &cDeepMimicCore::IsDone)
        .def("SetPlaybackSpeed", &cDeepMimicCore::SetPlaybackSpeed).def("SetUpdatesPerSec", &cDeepMimicCore::SetUpdatesPerSec)
        .def("GetWinWidth", &cDeepMimicCore::GetWinWidth).def("GetWinHeight", &cDeepMimicCore::GetWinHeight)
        .def("GetNumUpdateSubsteps", &cDeepMimicCore::GetNumUpdateSubsteps).def("IsRLScene", &cDeepMimicCore::IsRLScene)
        .def("GetNumAgents", &cDeepMimicCore::GetNumAgents).def("NeedNewAction", &cDeepMimicCore::NeedNewAction)
        .def("RecordState", &cDeepMimicCore::RecordState).def("RecordGoal", &cDeepMimicCore::RecordGoal).def("SetAction", &cDeepMimicCore::SetAction)
        .def("LogVal", &cDeepMimicCore::LogVal).def("GetActionSpace", &cDeepMimicCore::GetActionSpace).def("GetStateSize", &cDeepMimicCore::GetStateSize)
        .def("GetGoalSize", &cDeepMimicCore::GetGoalSize).def("GetActionSize", &cDeepMimicCore::GetActionSize).def("GetNumActions", &cDeepMimicCore::GetNumActions)
        .def("BuildStateOffset", &cDeepMimicCore::BuildStateOffset).def("BuildStateScale", &cDeepMimicCore::BuildStateScale)
        .def("BuildGoalOffset", &cDeepMimicCore::BuildGoalOffset).def("BuildGoalScale", &cDeepMimicCore::BuildGoalScale)
        .def("BuildActionOffset", &cDeepMimicCore::BuildActionOffset).def("BuildActionScale", &cDeepMimicCore::BuildActionScale)
        .def("BuildActionBoundMin", &cDeepMimicCore::BuildActionBoundMin).def("BuildActionBoundMax", &cDeepMimicCore::BuildActionBoundMax)
        .def("BuildStateNormGroups", &cDeepMimicCore::BuildStateNormGroups).def("BuildGoalNormGroups", &cDeepMimicCore::BuildGoalNormGroups)
        .def("CalcReward", &cDeepMimicCore::CalcReward).def("GetRewardMin", &cDeepMimicCore::GetRewardMin).def("GetRewardMax", &cDeepMimicCore::GetRewardMax)
        .def("GetRewardFail", &cDeepMimicCore::GetRewardFail).def("GetRewardSucc", &cDeepMimicCore::GetRewardSucc)
        .def("EnableAMPTaskReward", &cDeepMimicCore::EnableAMPTaskReward).def("GetAMPObsSize", &cDeepMimicCore::GetAMPObsSize)
        .def("GetAMPObsOffset", &cDeepMimicCore::GetAMPObsOffset).def("GetAMPObsScale", &cDeepMimicCore::GetAMPObsScale)
        .def("GetAMPObsNormGroup", &cDeepMimicCore::GetAMPObsNormGroup).def("RecordAMPObsExpert", &cDeepMimicCore::RecordAMPObsExpert)
        .def("RecordAMPObsAgent", &cDeepMimicCore::RecordAMPObsAgent).def("IsEpisodeEnd", &cDeepMimicCore::IsEpisodeEnd)
        .def("CheckValidEpisode", &cDeepMimicCore::CheckValidEpisode).def("CheckTerminate", &cDeepMimicCore::CheckTerminate)
        .def("SetMode", &cDeepMimicCore::SetMode).def("SetSampleCount", &cDeepMimicCore::SetSampleCount).def("GetNumEnvs", &cDeepMimicCore::GetNumEnvs);
}
