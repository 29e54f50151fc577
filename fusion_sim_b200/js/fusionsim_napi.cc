// fusionsim_napi.cc -- thin N-API addon over include/fusionsim.h (SOURCE ONLY: this image has no
// Node.js and no node_api.h, so it is not built or run here; tests/test_abi.py type-checks it against a
// declarations-only stand-in for napi.h, tests/stubs/napi.h).
//
// Build where Node exists (one command):
//   g++ -std=c++17 -shared -fPIC -I$(node -p "require('node-addon-api').include_dir")
//       -I../../include fusionsim_napi.cc -L../csrc -lfusionsim -Wl,-rpath,'$ORIGIN/../csrc'
//       -o fusionsim.node
// It owns nothing but the fsim_sim handle; typed arrays are passed straight through as host
// pointers (the library copies during the call, as gl.texImage2D does: utilities.js:585-594).
#include <napi.h>

#include <string>

#include "fusionsim.h"

namespace {

void check(Napi::Env env, int rc)
{
    if (rc != FSIM_OK) throw Napi::Error::New(env, fsim_last_error());  // reference: throw new Error(...)
}

class Sim : public Napi::ObjectWrap<Sim> {
  public:
    static Napi::Function Init(Napi::Env env)
    {
        return DefineClass(env, "Sim", {
            InstanceMethod("setArray", &Sim::SetArray),
            InstanceMethod("setSourcePdf", &Sim::SetSourcePdf),
            InstanceMethod("addCurrentLoop", &Sim::AddCurrentLoop),
            InstanceMethod("addCurrentZ", &Sim::AddCurrentZ),
            InstanceMethod("addBZ", &Sim::AddBZ),
            InstanceMethod("addBTheta", &Sim::AddBTheta),
            InstanceMethod("addSpindleCuspPlasmaField", &Sim::AddSpindle),
            InstanceMethod("precalc", &Sim::Precalc),
            InstanceMethod("step", &Sim::Step),
            InstanceMethod("density", &Sim::Density),
            InstanceMethod("solveFields", &Sim::SolveFields),
            InstanceMethod("runFrames", &Sim::RunFrames),
            InstanceMethod("halfStep", &Sim::HalfStep),
            InstanceMethod("emInit", &Sim::EmInit),
            InstanceMethod("emSet", &Sim::EmSet),
            InstanceMethod("emGet", &Sim::EmGet),
            InstanceMethod("emStep", &Sim::EmStep),
            InstanceMethod("setState", &Sim::SetState),
            InstanceMethod("setField", &Sim::SetField),
            InstanceMethod("render", &Sim::Render),
            InstanceMethod("getArray", &Sim::GetArray),
            InstanceMethod("destroy", &Sim::Destroy),
        });
    }

    explicit Sim(const Napi::CallbackInfo &info) : Napi::ObjectWrap<Sim>(info)
    {
        Napi::Object o = info[0].As<Napi::Object>();
        fsim_spec s{};
        auto num = [&](const char *k) { return o.Get(k).As<Napi::Number>().DoubleValue(); };
        s.radius = num("radius"); s.height = num("height");
        s.nr = (int64_t)num("nr"); s.nz = (int64_t)num("nz");
        s.dt = num("dt"); s.nparticles = (int64_t)num("nparticles");
        s.particle_mass = num("particle_mass"); s.particle_charge = num("particle_charge");
        if (o.Has("precision")) s.precision = o.Get("precision").As<Napi::Number>().Int32Value();
        if (o.Has("device")) s.device = o.Get("device").As<Napi::Number>().Int32Value();
        if (o.Has("flags")) s.flags = o.Get("flags").As<Napi::Number>().Uint32Value();
        check(info.Env(), fsim_create(&s, &sim_));
    }
    ~Sim() { fsim_destroy(sim_); }

  private:
    fsim_sim *sim_ = nullptr;

    // setArray(name, Float64Array): E, B, position, velocity, sink_mask, rand, entropy, inv_cdf
    Napi::Value SetArray(const Napi::CallbackInfo &info)
    {
        std::string name = info[0].As<Napi::String>();
        const double *p = info[1].As<Napi::Float64Array>().Data();
        int rc = FSIM_ERR_INVALID;
        if (name == "E") rc = fsim_set_E(sim_, p);
        else if (name == "B") rc = fsim_set_B(sim_, p);
        else if (name == "position") rc = fsim_set_position(sim_, p);
        else if (name == "velocity") rc = fsim_set_velocity(sim_, p);
        else if (name == "sink_mask") rc = fsim_set_sink_mask(sim_, p);
        else if (name == "rand") rc = fsim_set_rand(sim_, p);
        else if (name == "entropy") rc = fsim_set_entropy(sim_, p);
        else if (name == "inv_cdf") rc = fsim_set_inv_cdf(sim_, p);
        check(info.Env(), rc);
        return info.Env().Undefined();
    }
    Napi::Value SetSourcePdf(const Napi::CallbackInfo &info)
    {
        check(info.Env(), fsim_set_source_pdf(sim_, info[0].As<Napi::Float64Array>().Data(),
                                              info[1].As<Napi::Number>().Int64Value(),
                                              info[2].As<Napi::Number>().Int64Value()));
        return info.Env().Undefined();
    }
    Napi::Value AddCurrentLoop(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_add_current_loop(sim_, i[0].As<Napi::Number>(), i[1].As<Napi::Number>(),
                                             i[2].As<Napi::Number>()));
        return i.Env().Undefined();
    }
    Napi::Value AddCurrentZ(const Napi::CallbackInfo &i) { check(i.Env(), fsim_add_current_z(sim_, i[0].As<Napi::Number>())); return i.Env().Undefined(); }
    Napi::Value AddBZ(const Napi::CallbackInfo &i) { check(i.Env(), fsim_add_bz(sim_, i[0].As<Napi::Number>())); return i.Env().Undefined(); }
    Napi::Value AddBTheta(const Napi::CallbackInfo &i) { check(i.Env(), fsim_add_btheta(sim_, i[0].As<Napi::Number>())); return i.Env().Undefined(); }
    // addSpindleCuspPlasmaField(r, B_c, beta_c = 1): empic.js:1369; the boundary solve specified in include/fusionsim.h
    Napi::Value AddSpindle(const Napi::CallbackInfo &i)
    {
        const double beta = i.Length() > 2 && i[2].IsNumber() ? i[2].As<Napi::Number>().DoubleValue() : 1.0;
        check(i.Env(), fsim_add_spindle_cusp_plasma_field(sim_, i[0].As<Napi::Number>(), i[1].As<Napi::Number>(), beta));
        return i.Env().Undefined();
    }
    Napi::Value Precalc(const Napi::CallbackInfo &i) { check(i.Env(), fsim_precalc(sim_)); return i.Env().Undefined(); }
    Napi::Value Step(const Napi::CallbackInfo &i) { check(i.Env(), fsim_step(sim_)); return i.Env().Undefined(); }
    Napi::Value Density(const Napi::CallbackInfo &i) { check(i.Env(), fsim_density(sim_)); return i.Env().Undefined(); }
    // checkpoint restore (extension): setState(position4 | null, velocity3 | null, rand4 | null) as Float64Arrays
    Napi::Value SetState(const Napi::CallbackInfo &i)
    {
        const double *p[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < 3; ++k)
            if (i[k].IsTypedArray()) p[k] = i[k].As<Napi::Float64Array>().Data();
        check(i.Env(), fsim_set_state(sim_, p[0], p[1], p[2]));
        return i.Env().Undefined();
    }
    // setField("moments01_avg" | "phi", Float64Array)
    Napi::Value SetField(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_set_field(sim_, i[0].As<Napi::String>().Utf8Value().c_str(), i[1].As<Napi::Float64Array>().Data()));
        return i.Env().Undefined();
    }
    // EXTENSION (no reference counterpart): solveFields(macro_weight, sweeps, omega, source)
    Napi::Value SolveFields(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_solve_fields(sim_, i[0].As<Napi::Number>().DoubleValue(), i[1].As<Napi::Number>().Int32Value(),
                                         i[2].As<Napi::Number>().DoubleValue(), i[3].As<Napi::Number>().Int32Value()));
        return i.Env().Undefined();
    }
    // runFrames(n): n iterations of the page loop (step, density with its canvas draws); launch-bound scenes replay a
    // captured CUDA graph (include/fusionsim.h, fsim_run_frames)
    Napi::Value RunFrames(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_run_frames(sim_, i[0].As<Napi::Number>().Int64Value()));
        return i.Env().Undefined();
    }
    // EXTENSION (no reference counterpart): electromagnetic update on a Yee mesh -- emInit(), emSet(name, Float64Array),
    // emGet(name, Float64Array out), emStep(macro_weight, with_current); a frame is halfStep() + density() + emStep()
    Napi::Value HalfStep(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_half_step(sim_));
        return i.Env().Undefined();
    }
    Napi::Value EmInit(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_em_init(sim_));
        return i.Env().Undefined();
    }
    Napi::Value EmSet(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_em_set(sim_, i[0].As<Napi::String>().Utf8Value().c_str(), i[1].As<Napi::Float64Array>().Data()));
        return i.Env().Undefined();
    }
    Napi::Value EmGet(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_em_get(sim_, i[0].As<Napi::String>().Utf8Value().c_str(), i[1].As<Napi::Float64Array>().Data()));
        return i.Env().Undefined();
    }
    Napi::Value EmStep(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_em_step(sim_, i[0].As<Napi::Number>().DoubleValue(), i[1].ToBoolean().Value() ? 1 : 0));
        return i.Env().Undefined();
    }
    Napi::Value Render(const Napi::CallbackInfo &i)
    {
        check(i.Env(), fsim_render_rgba8(sim_, i[0].As<Napi::Uint8Array>().Data()));
        return i.Env().Undefined();
    }
    // getArray(name, Float64Array out): position, velocity, rand, or a field name
    Napi::Value GetArray(const Napi::CallbackInfo &info)
    {
        std::string name = info[0].As<Napi::String>();
        double *p = info[1].As<Napi::Float64Array>().Data();
        int rc;
        if (name == "position") rc = fsim_get_position(sim_, p);
        else if (name == "velocity") rc = fsim_get_velocity(sim_, p);
        else if (name == "rand") rc = fsim_get_rand(sim_, p);
        else rc = fsim_get_field(sim_, name.c_str(), p);
        check(info.Env(), rc);
        return info.Env().Undefined();
    }
    Napi::Value Destroy(const Napi::CallbackInfo &i) { fsim_destroy(sim_); sim_ = nullptr; return i.Env().Undefined(); }
};

Napi::Object InitAll(Napi::Env env, Napi::Object exports)
{
    exports.Set("Sim", Sim::Init(env));
    return exports;
}

}  // namespace

NODE_API_MODULE(fusionsim, InitAll)
