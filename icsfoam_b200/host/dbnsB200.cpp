// dbnsB200 — standalone driver: dbnsFoam's time / pseudo-time loop (dbnsFoam.C:88-151, outerLoop.H, updateFields.H,
// pseudotimeControl.C:72-101,159-241) over the icsb200 C ABI, reading an OpenFOAM case directory verbatim:
//   constant/polyMesh/{points,faces,owner,neighbour,boundary}, constant/thermophysicalProperties,
//   system/{controlDict,fvSchemes,fvSolution}, 0/{p,U,T} (uniform internalField + boundaryField types).
// It exists because OpenFOAM v2112 is not available in this image (the OpenFOAM adapter is shown in INTEGRATION.md);
// the dictionary keys are the reference's (SURVEY.md §5 "Config / flags").
//
// Fields may be `uniform` or `nonuniform List<scalar|vector>` (internalField of a written time directory; value / p0 / T0 /
// inletValue profiles on patches).  -writeFields writes p, U, T of the last step as <caseDir>/<time>/{p,U,T} (ascii, boundary
// patches as `calculated` with their values; copy the internalField into 0/ to restart).  -parseOnly stops after reading the case.
//
//   dbnsB200 <caseDir> [-maxSteps N] [-device D] [-writeFields] [-writeFlux] [-parseOnly]
// -writeFlux (implies -writeFields) adds rho and the face fluxes phi, phiUp, phiEp of the LAST outer iteration (evaluated, as in
// outerLoop.H:51-57, from the state that iteration started from — what dbnsFoam's AUTO_WRITE phi holds at runTime.write()), so
// that a time directory of the reference run on an OpenFOAM machine can be compared file by file (tools/foamdiff.py).
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <iostream>

#include "icsfoamB200.H"

using namespace icsfoamB200;

namespace {

// ---- meshtools C API (libicsmesh.so), loaded at run time ----
struct MeshLib {
    void* h = nullptr;
    void* (*read_polymesh)(const char*) = nullptr;
    const char* (*error)(void*) = nullptr;
    void (*sizes)(void*, int*) = nullptr;
    void (*patch)(void*, int, int*, char*) = nullptr;
    void (*arrays)(void*, int*, int*, double*, double*, double*, double*, double*, double*, double*, double*) = nullptr;
    void (*free_)(void*) = nullptr;
    explicit MeshLib(const std::string& path)
    {
        h = dlopen(path.c_str(), RTLD_NOW);
        if (!h) throw FatalError(std::string("cannot load ") + path + ": " + dlerror());
        read_polymesh = (decltype(read_polymesh))dlsym(h, "icsmesh_read_polymesh");
        error = (decltype(error))dlsym(h, "icsmesh_error");
        sizes = (decltype(sizes))dlsym(h, "icsmesh_sizes");
        patch = (decltype(patch))dlsym(h, "icsmesh_patch");
        arrays = (decltype(arrays))dlsym(h, "icsmesh_arrays");
        free_ = (decltype(free_))dlsym(h, "icsmesh_free");
    }
};

struct BcSpec { int kind; std::vector<double> prm; int nRows = 0; /* > 0: prm holds nRows x (prm.size()/nRows) non-uniform rows */ };

std::vector<double> numbers(const std::vector<std::string>& toks)
{
    std::vector<double> v;
    for (auto& t : toks) {
        char* end = nullptr;
        double x = std::strtod(t.c_str(), &end);
        if (end && *end == 0 && !t.empty()) v.push_back(x);
    }
    return v;
}

// `uniform v...` -> n copies of the nc components; `nonuniform List<...> n ( ... )` -> the n x nc listed values
std::vector<double> fieldValues(const std::vector<std::string>& toks, int nc, int n, const std::string& what)
{
    std::vector<double> nums = numbers(toks), out((size_t)nc * n);
    if (!toks.empty() && toks[0] == "nonuniform") {
        if (nums.empty() || (long)nums[0] != n || (long)nums.size() != 1 + (long)nc * n)
            throw FatalError("nonuniform " + what + ": expected " + std::to_string(n) + " x " + std::to_string(nc) + " values");
        std::copy(nums.begin() + 1, nums.end(), out.begin());
    } else {
        if ((int)nums.size() != nc) throw FatalError("uniform " + what + ": expected " + std::to_string(nc) + " components");
        for (int i = 0; i < n; i++) for (int k = 0; k < nc; k++) out[(size_t)nc * i + k] = nums[k];
    }
    return out;
}

void writeField(const std::string& path, const char* cls, const char* name, const char* dims, int nc, const std::vector<double>& cells,
                const std::vector<std::string>& patchNames, const std::vector<icsb200_patch>& patches, int F, const std::vector<double>& bnd)
{
    std::ofstream os(path);
    if (!os) throw FatalError("cannot write " + path);
    os.precision(17);
    auto list = [&](const double* v, size_t n) {
        os << "nonuniform List<" << (nc == 1 ? "scalar" : "vector") << "> " << n << "\n(\n";
        for (size_t i = 0; i < n; i++) {
            if (nc == 1) os << v[i] << "\n";
            else os << "(" << v[3 * i] << " " << v[3 * i + 1] << " " << v[3 * i + 2] << ")\n";
        }
        os << ")";
    };
    os << "FoamFile\n{\n    version 2.0;\n    format ascii;\n    class " << cls << ";\n    object " << name << ";\n}\n\ndimensions " << dims
       << ";\n\ninternalField ";
    list(cells.data(), cells.size() / nc);
    os << ";\n\nboundaryField\n{\n";
    for (size_t pi = 0; pi < patches.size(); pi++) {
        os << "    " << patchNames[pi] << "\n    {\n";
        if (patches[pi].kind == ICSB200_EMPTY) os << "        type empty;\n";
        else if (patches[pi].kind == ICSB200_CYCLIC) os << "        type cyclic;\n";
        else {
            os << "        type calculated;\n        value ";
            list(bnd.data() + (size_t)nc * (patches[pi].start - F), patches[pi].size);
            os << ";\n";
        }
        os << "    }\n";
    }
    os << "}\n";
}

// fvPatchField type name -> ICSB200_BC_* + parameters (the BC set of the five tutorials)
BcSpec bcFromDict(const dictionary& d, int field, const std::vector<double>& Uinf, const dictionary& top, int nFaces)
{
    const std::string type = d.word("type");
    BcSpec b{ICSB200_BC_ZEROGRADIENT, {}};
    // `$internalField`-style macros refer to entries of the enclosing field file
    auto val = [&](const char* k) {
        std::vector<std::string> toks = d.lookup(k);
        if (!toks.empty() && toks[0].size() > 1 && toks[0][0] == '$') toks = top.lookup(toks[0].substr(1));
        return numbers(toks);
    };
    if (type == "zeroGradient") b.kind = ICSB200_BC_ZEROGRADIENT;
    else if (type == "fixedValue") { b.kind = ICSB200_BC_FIXEDVALUE; b.prm = val("value"); }
    else if (type == "slip" || type == "symmetryPlane" || type == "symmetry") b.kind = ICSB200_BC_SLIP;
    else if (type == "empty") b.kind = ICSB200_BC_EMPTY;
    else if (type == "inletOutlet") { b.kind = ICSB200_BC_INLETOUTLET; b.prm = val("inletValue"); }
    else if (type == "freestream") { b.kind = ICSB200_BC_INLETOUTLET; b.prm = val("freestreamValue"); }
    else if (type == "totalPressure") { b.kind = ICSB200_BC_TOTALPRESSURE; b.prm = {val("p0").at(0), d.get<double>("gamma")}; }
    else if (type == "totalTemperature") { b.kind = ICSB200_BC_TOTALTEMPERATURE; b.prm = {val("T0").at(0), d.get<double>("gamma")}; }
    else if (type == "pressureInletOutletVelocity") {
        b.kind = ICSB200_BC_PRESSUREINLETOUTLETVELOCITY;
        b.prm = d.found("tangentialVelocity") ? val("tangentialVelocity") : std::vector<double>{0, 0, 0};
    } else if (type == "freestreamPressure") {
        b.kind = ICSB200_BC_FREESTREAMPRESSURE;
        b.prm = {val("freestreamValue").at(0), Uinf.at(0), Uinf.at(1), Uinf.at(2)};
    } else if (type == "noSlip") { b.kind = ICSB200_BC_FIXEDVALUE; b.prm = {0, 0, 0}; }
    else if (type == "cyclic" || type == "processor") b.kind = ICSB200_BC_COUPLED;
    else throw FatalError("unsupported patch field type " + type + " (supported: zeroGradient fixedValue slip symmetryPlane empty inletOutlet "
                          "freestream totalPressure totalTemperature pressureInletOutletVelocity freestreamPressure noSlip)");
    // non-uniform entries: value (fixedValue), p0 / T0 (+ gamma), inletValue — one row per face
    const int nc = field == 1 ? 3 : 1;
    const char* key = type == "fixedValue" ? "value" : type == "totalPressure" ? "p0" : type == "totalTemperature" ? "T0"
                      : type == "inletOutlet" ? "inletValue" : nullptr;
    if (key && d.found(key) && !d.lookup(key).empty() && d.lookup(key)[0] == "nonuniform") {
        const std::vector<double> v = fieldValues(d.lookup(key), nc, nFaces, std::string(key));
        const bool withGamma = type == "totalPressure" || type == "totalTemperature";
        const int np = nc + (withGamma ? 1 : 0);
        b.prm.assign((size_t)np * nFaces, 0.0);
        for (int i = 0; i < nFaces; i++) {
            for (int k = 0; k < nc; k++) b.prm[(size_t)np * i + k] = v[(size_t)nc * i + k];
            if (withGamma) b.prm[(size_t)np * i + nc] = d.get<double>("gamma");
        }
        b.nRows = nFaces;
    }
    return b;
}

}  // namespace

int main(int argc, char** argv)
{
    try {
        if (argc < 2) { std::fprintf(stderr, "usage: dbnsB200 <caseDir> [-maxSteps N] [-device D]\n"); return 2; }
        const std::string caseDir = argv[1];
        int maxSteps = 1 << 30, device = 0;
        bool writeFields = false, writeFlux = false, parseOnly = false;
        for (int i = 2; i < argc; i++) {
            if (!std::strcmp(argv[i], "-maxSteps") && i + 1 < argc) maxSteps = std::atoi(argv[++i]);
            else if (!std::strcmp(argv[i], "-device") && i + 1 < argc) device = std::atoi(argv[++i]);
            else if (!std::strcmp(argv[i], "-writeFields")) writeFields = true;
            else if (!std::strcmp(argv[i], "-writeFlux")) writeFields = writeFlux = true;
            else if (!std::strcmp(argv[i], "-parseOnly")) parseOnly = true;
        }
        const char* libEnv = std::getenv("ICSMESH_LIB");
        MeshLib ml(libEnv ? libEnv : "libicsmesh.so");

        // ---- dictionaries
        const dictionary controlDict = dictionary::fromFile(caseDir + "/system/controlDict");
        const dictionary fvSchemes = dictionary::fromFile(caseDir + "/system/fvSchemes");
        const dictionary fvSolution = dictionary::fromFile(caseDir + "/system/fvSolution");
        const dictionary thermoDict = dictionary::fromFile(caseDir + "/constant/thermophysicalProperties");
        const dictionary pDict = dictionary::fromFile(caseDir + "/0/p"), UDict = dictionary::fromFile(caseDir + "/0/U"),
                         TDict = dictionary::fromFile(caseDir + "/0/T");

        // ---- mesh
        void* mh = ml.read_polymesh((caseDir + "/constant/polyMesh").c_str());
        if (ml.error(mh)[0]) throw FatalError(ml.error(mh));
        int sz[7];
        ml.sizes(mh, sz);
        const int N = sz[0], F = sz[1], FT = sz[2], nP = sz[3];
        std::vector<int> owner(FT), neighbour(F);
        std::vector<double> Sf(3 * (size_t)FT), Cf(3 * (size_t)FT), magSf(FT), w(FT), dc(FT), nodc(FT), C(3 * (size_t)N), V(N);
        ml.arrays(mh, owner.data(), neighbour.data(), Sf.data(), Cf.data(), magSf.data(), w.data(), dc.data(), nodc.data(), C.data(), V.data());
        std::vector<icsb200_patch> patches(nP);
        std::vector<std::string> patchNames(nP);
        for (int i = 0; i < nP; i++) {
            int o[5];
            char name[64];
            ml.patch(mh, i, o, name);
            patchNames[i] = name;
            patches[i] = icsb200_patch{o[0], o[1], o[2], o[3], o[4], {1, 0, 0, 0, 1, 0, 0, 0, 1}};
        }
        std::cout << "Mesh: " << N << " cells, " << F << " internal faces, " << nP << " patches\n";

        // ---- initial fields and boundary conditions (parsed before the device is touched)
        std::vector<double> p = fieldValues(pDict.lookup("internalField"), 1, N, "internalField of p");
        std::vector<double> U = fieldValues(UDict.lookup("internalField"), 3, N, "internalField of U");
        std::vector<double> T = fieldValues(TDict.lookup("internalField"), 1, N, "internalField of T");
        const dictionary *top[3] = {&pDict, &UDict, &TDict};
        const dictionary *bf[3] = {&pDict.subDict("boundaryField"), &UDict.subDict("boundaryField"), &TDict.subDict("boundaryField")};
        std::vector<BcSpec> bcs((size_t)3 * nP);
        int nNonuniform = 0;
        for (int pi = 0; pi < nP; pi++)
            for (int fld = 0; fld < 3; fld++) {
                if (!bf[fld]->isDict(patchNames[pi])) throw FatalError("patch " + patchNames[pi] + " missing in boundaryField");
                // freestreamPressureFvPatchScalarField blends with the velocity of the U patch field on the same patch: its
                // freestreamValue (macros resolved against 0/U), not the internalField
                std::vector<double> Ufs{0, 0, 0};
                if (fld == 0 && bf[0]->subDict(patchNames[pi]).word("type") == "freestreamPressure") {
                    const dictionary& up = bf[1]->subDict(patchNames[pi]);
                    if (up.word("type") != "freestream") throw FatalError("patch " + patchNames[pi] + ": freestreamPressure needs a freestream U patch field");
                    if (bf[0]->subDict(patchNames[pi]).getSwitch("supersonic", false)) throw FatalError("patch " + patchNames[pi] + ": freestreamPressure with supersonic true is not supported");
                    std::vector<std::string> toks = up.lookup("freestreamValue");
                    if (!toks.empty() && toks[0].size() > 1 && toks[0][0] == '$') toks = UDict.lookup(toks[0].substr(1));
                    Ufs = numbers(toks);
                    if (Ufs.size() < 3) throw FatalError("patch " + patchNames[pi] + ": freestreamValue of U must be a uniform vector");
                }
                bcs[(size_t)3 * pi + fld] = bcFromDict(bf[fld]->subDict(patchNames[pi]), fld, Ufs, *top[fld], patches[pi].size);
                nNonuniform += bcs[(size_t)3 * pi + fld].nRows > 0;
            }
        // ---- thermo: hePsiThermo<pureMixture<constTransport<hConst<perfectGas>>>>, sensibleInternalEnergy (createFields.H:17-35)
        const dictionary& mix = thermoDict.subDict("mixture");
        const double W = mix.subDict("specie").get<double>("molWeight"), Cp = mix.subDict("thermodynamics").get<double>("Cp");
        const double mu = mix.subDict("transport").get<double>("mu"), Pr = mix.subDict("transport").get<double>("Pr");
        // hConstThermo (v2112): Hs = Cp (T - Tref) + Hsref with Tref defaulting to Tstd = 298.15 K when the entry is absent, so
        // e = Cv T - Cp Tref + Hsref.  The device thermo is e = Cv T: every shipped tutorial says `Tref 0`; anything else is refused.
        if (mix.subDict("thermodynamics").getOrDefault<double>("Tref", 298.15) != 0.0 || mix.subDict("thermodynamics").getOrDefault<double>("Hsref", 0.0) != 0.0)
            throw FatalError("thermodynamics: only `Tref 0` (stated explicitly; OpenFOAM defaults to Tstd) with Hsref 0 is supported");
        // ---- solver controls (fvSolution/flowSolver): read before -parseOnly returns so that dictionary errors surface without a GPU
        const icsb200_solver_controls ctl = coupledMatrix::controlsFromDict(fvSolution.subDict("flowSolver"));
        // ---- time scheme and run controls this driver does not implement are refused, not silently replaced (also before -parseOnly returns)
        const std::vector<std::string>& ddt = fvSchemes.subDict("ddtSchemes").lookup("default");
        if (ddt.at(0) != "dualTime") throw FatalError("ddtSchemes default must be 'dualTime rPseudoDeltaT <inner>' (dualTimeDdtScheme.H:103)");
        const std::string inner = ddt.back();
        const bool steadyState = inner == "steadyState";
        if (!steadyState && inner != "Euler" && inner != "backward") throw FatalError("inner ddt scheme '" + inner + "' is not supported (steadyState Euler backward)");
        if (fvSolution.subDict("pseudoTime").getSwitch("resetPseudo", false)) throw FatalError("pseudoTime/resetPseudo true (beginTimeStep.H) is not supported");
        if (controlDict.getSwitch("adjustTimeStep", false)) throw FatalError("controlDict adjustTimeStep yes is not supported");
        if (parseOnly) {
            double pmin = 1e300, pmax = -1e300;
            for (double v : p) { pmin = std::min(pmin, v); pmax = std::max(pmax, v); }
            std::cout << "parse ok: p in [" << pmin << ", " << pmax << "], " << nNonuniform << " non-uniform patch entries\nEnd\n";
            ml.free_(mh);
            return 0;
        }

        icsb200_ctx* ctx = nullptr;
        check(nullptr, icsb200_create(&ctx, device, nullptr, 0, 1), "icsb200_create (needs a B200; there is no CPU fallback)");
        check(ctx, icsb200_mesh_set(ctx, N, F, FT, owner.data(), neighbour.data(), Sf.data(), magSf.data(), w.data(), dc.data(), nodc.data(),
                                    C.data(), V.data(), Cf.data(), nP, patches.data(), sz + 4), "mesh_set");

        check(ctx, icsb200_thermo_set(ctx, 8314.46261815324 / W, Cp, mu, Pr), "thermo_set");
        if (mu > 0) std::cout << "Viscous analysis detected: laminar viscous residual + Lax-Friedrichs viscous Jacobian (turbulence model not on the device)\n";

        // ---- schemes (fvSchemes / fvSolution pseudoTime; initialise.H:39-71, beginTimeStep.H:8-45, updateFields.H:11-35)
        auto flux = convectiveFluxScheme::New(ctx, fvSchemes);
        const dictionary& interp = fvSchemes.subDict("interpolationSchemes");
        const dictionary& pseudo = fvSolution.subDict("pseudoTime");
        icsb200_schemes sch{};
        sch.flux_scheme = flux->id();
        sch.limiter_rho = limiterId(interp.word("reconstruct(rho)"));
        sch.limiter_U = limiterId(interp.word("reconstruct(U)"));
        sch.limiter_T = limiterId(interp.word("reconstruct(T)"));
        const dictionary& cfs = fvSchemes.subDict("convectiveFluxScheme");
        sch.low_mach_ausm = cfs.getSwitch("lowMachAusm", true);
        sch.entropy_fix_coeff = cfs.getOrDefault<double>("entropyFixCoeff", 0.05);
        sch.ddt_scheme = steadyState ? ICSB200_DDT_STEADY : inner == "Euler" ? ICSB200_DDT_EULER : ICSB200_DDT_BACKWARD;
        sch.delta_t = controlDict.get<double>("deltaT");
        sch.local_timestepping = pseudo.getSwitch("localTimestepping", true);
        sch.local_timestepping_bounding = pseudo.getSwitch("localTimesteppingBounding", true);
        sch.local_timestepping_lower_bound = std::min(std::max(pseudo.getOrDefault<double>("localTimesteppingLowerBound", 0.95), 0.0), 0.99);
        sch.pseudo_co_num = pseudo.getOrDefault<double>("pseudoCoNum", 1.0);
        sch.pseudo_co_num_min = pseudo.getOrDefault<double>("pseudoCoNumMin", 0.1);  // 'pseudoCoMin' is never read (Q11)
        sch.pseudo_co_num_max = pseudo.getOrDefault<double>("pseudoCoNumMax", 25.0);
        sch.pseudo_co_num_max_incr = pseudo.getOrDefault<double>("pseudoCoNumMaxIncreaseFactor", 1.25);
        sch.pseudo_co_num_min_decr = pseudo.getOrDefault<double>("pseudoCoNumMinDecreaseFactor", 0.1);
        sch.rho_min = pseudo.getOrDefault<double>("rhoMin", -1e15);
        sch.T_min = pseudo.getOrDefault<double>("TMin", 1e-15);
        sch.T_max = pseudo.getOrDefault<double>("TMax", 1e15);
        sch.viscous_full_jacobian = viscousFluxScheme(ctx, fvSchemes).fullJacobian();
        check(ctx, icsb200_schemes_set(ctx, &sch), "schemes_set");
        std::cout << (steadyState ? "Steady-state analysis detected\n" : "Transient analysis detected\n");

        // ---- boundary conditions and initial fields
        for (int pi = 0; pi < nP; pi++)
            for (int fld = 0; fld < 3; fld++) {
                const BcSpec& b = bcs[(size_t)3 * pi + fld];
                if (b.kind == ICSB200_BC_EMPTY || b.kind == ICSB200_BC_COUPLED) continue;
                if (b.nRows > 0) check(ctx, icsb200_bc_set_nonuniform(ctx, pi, fld, b.kind, b.prm.data(), (int)(b.prm.size() / b.nRows)), "bc_set_nonuniform");
                else check(ctx, icsb200_bc_set(ctx, pi, fld, b.kind, b.prm.data(), (int)b.prm.size()), "bc_set");
            }
        check(ctx, icsb200_state_set(ctx, p.data(), U.data(), T.data()), "state_set");

        // ---- solver controls (fvSolution/flowSolver) and pseudo-time control (pseudotimeControl.C:42-69)
        coupledMatrix eqSystem(ctx);
        (void)eqSystem;
        const int nCorrOuter = pseudo.getOrDefault<int>("nPseudoCorr", 20), nCorrOuterMin = pseudo.getOrDefault<int>("nPseudoCorrMin", 1);
        const double pseudoTol = pseudo.get<double>("pseudoTol"), pseudoTolRel = pseudo.get<double>("pseudoTolRel");
        const double endTime = controlDict.get<double>("endTime"), deltaT = controlDict.get<double>("deltaT");
        double time = controlDict.getOrDefault<double>("startTime", 0.0);

        // ---- time loop (dbnsFoam.C:88-151)
        std::vector<double> phiH(writeFlux ? FT : 0), phiUpH(writeFlux ? 3 * (size_t)FT : 0), phiEpH(writeFlux ? FT : 0);
        pseudotimeControl solnControl(steadyState, nCorrOuter, nCorrOuterMin, pseudoTol, pseudoTolRel, std::cout);
        int step = 0;
        while (time < endTime - 1e-12 * std::fabs(endTime) && step < maxSteps) {
            time += deltaT;   // runTime++ (steady runs count pseudo iterations in units of controlDict's deltaT, too)
            step++;
            std::cout << "Time = " << time << "\n\n";
            if (!steadyState) check(ctx, icsb200_new_time_step(ctx), "new_time_step");   // beginTimeStep.H
            bool notFinished;
            while ((notFinished = solnControl.loop())) {
                icsb200_residuals r;
                if (writeFlux) check(ctx, icsb200_calc_flux(ctx, phiH.data(), phiUpH.data(), phiEpH.data()), "calc_flux");
                check(ctx, icsb200_iterate_dev(ctx, &ctl, &r), "outerLoop");   // outerLoop.H + updateFields.H on the device
                const residualsIO residuals(r);
                solnControl.setResidual(residuals);   // outerLoop.H:97
                std::cout << "GMRES : Solving for (  rhoIncr rhoEIncr rhoUIncr ) \n";
                residuals.print(std::cout);
                if (steadyState) break;
            }
            if (steadyState && !notFinished) break;   // runTime.writeAndEnd()
        }
        std::vector<double> rho(N);
        check(ctx, icsb200_state_get(ctx, rho.data(), nullptr, nullptr, nullptr, nullptr, nullptr), "state_get");
        double rmin = 1e300, rmax = -1e300;
        for (double v : rho) { rmin = std::min(rmin, v); rmax = std::max(rmax, v); }
        if (writeFields) {
            // time directory of the last step (dbnsFoam: runTime.write()): p, U, T with the boundary values as `calculated` patches
            const int NBf = FT - F;
            std::vector<double> pb(NBf), Ub(3 * (size_t)NBf), Tb(NBf);
            check(ctx, icsb200_state_get(ctx, nullptr, nullptr, nullptr, p.data(), U.data(), T.data()), "state_get");
            check(ctx, icsb200_boundary_get(ctx, nullptr, Ub.data(), pb.data(), Tb.data()), "boundary_get");
            std::ostringstream tn;
            tn << time;
            const std::string dir = caseDir + "/" + tn.str();
            if (std::system(("mkdir -p '" + dir + "'").c_str()) != 0) throw FatalError("cannot create " + dir);
            writeField(dir + "/p", "volScalarField", "p", "[1 -1 -2 0 0 0 0]", 1, p, patchNames, patches, F, pb);
            writeField(dir + "/U", "volVectorField", "U", "[0 1 -1 0 0 0 0]", 3, U, patchNames, patches, F, Ub);
            writeField(dir + "/T", "volScalarField", "T", "[0 0 0 1 0 0 0]", 1, T, patchNames, patches, F, Tb);
            if (writeFlux) {
                std::vector<double> rhob(NBf);
                check(ctx, icsb200_boundary_get(ctx, rhob.data(), nullptr, nullptr, nullptr), "boundary_get");
                writeField(dir + "/rho", "volScalarField", "rho", "[1 -3 0 0 0 0 0]", 1, rho, patchNames, patches, F, rhob);
                auto surf = [&](const char* name, const char* cls, const char* dims, int nc, const std::vector<double>& v) {
                    const std::vector<double> in(v.begin(), v.begin() + (size_t)nc * F), bd(v.begin() + (size_t)nc * F, v.end());
                    writeField(dir + "/" + name, cls, name, dims, nc, in, patchNames, patches, F, bd);
                };
                surf("phi", "surfaceScalarField", "[1 0 -1 0 0 0 0]", 1, phiH);
                surf("phiUp", "surfaceVectorField", "[1 1 -2 0 0 0 0]", 3, phiUpH);
                surf("phiEp", "surfaceScalarField", "[1 2 -3 0 0 0 0]", 1, phiEpH);
                // LDU arrays of the nine sub-blocks of the last outer iteration's coupledMatrix (coupledMatrix.H:399-421), one ASCII
                // list per array: <time>/eqSystem/<block>_<diag|upper|lower> (tools/openfoam_golden/compare_matrix.py)
                static const char* blockNames[9] = {"dSByS_0_0", "dSByS_0_1", "dSByS_1_0", "dSByS_1_1", "dSByV_0_0", "dSByV_1_0", "dVByS_0_0", "dVByS_0_1", "dVByV_0_0"};
                static const int blockNc[9] = {1, 1, 1, 1, 3, 3, 3, 3, 9};
                if (std::system(("mkdir -p '" + dir + "/eqSystem'").c_str()) != 0) throw FatalError("cannot create " + dir + "/eqSystem");
                for (int b = 0; b < 9; b++) {
                    const int nc = blockNc[b];
                    std::vector<double> dg((size_t)nc * N), up((size_t)nc * F), lw((size_t)nc * F);
                    check(ctx, icsb200_matrix_get_ldu(ctx, b, dg.data(), up.data(), lw.data()), "matrix_get_ldu");
                    auto dump = [&](const char* part, const std::vector<double>& v) {
                        std::ofstream os(dir + "/eqSystem/" + blockNames[b] + "_" + part);
                        os.precision(17);
                        const size_t n = v.size() / nc;
                        os << n << "\n(\n";
                        for (size_t i = 0; i < n; i++) {
                            if (nc == 1) os << v[i] << "\n";
                            else { os << "("; for (int k = 0; k < nc; k++) os << (k ? " " : "") << v[i * nc + k]; os << ")\n"; }
                        }
                        os << ")\n";
                    };
                    dump("diag", dg); dump("upper", up); dump("lower", lw);
                }
            }
            std::cout << "fields written to " << dir << "\n";
        }
        std::cout << "rho min/max: " << rmin << " " << rmax << "\nkernel launches: " << icsb200_launch_count(ctx) << "\nEnd\n";
        icsb200_destroy(ctx);
        ml.free_(mh);
        return 0;
    } catch (const FatalError& e) {
        std::cerr << "--> FOAM FATAL ERROR:\n" << e.what() << "\n";
        return 1;
    }
}
