// NOT compiled in this repository's image (no OpenFOAM) — see README.md.
#include "B200FluxScheme.H"
#include "addToRunTimeSelectionTable.H"

namespace Foam
{
    defineTemplateTypeNameAndDebugWithName(hllcB200FluxScheme, "HLLCB200", 0);
    defineTemplateTypeNameAndDebugWithName(roeB200FluxScheme, "ROEB200", 0);
    defineTemplateTypeNameAndDebugWithName(ausmPlusUpB200FluxScheme, "AUSMPlusUpB200", 0);
    defineTemplateTypeNameAndDebugWithName(rusanovB200FluxScheme, "RusanovB200", 0);   // no CPU counterpart in ICSFoam
    addToRunTimeSelectionTable(convectiveFluxScheme, hllcB200FluxScheme, dictionary);
    addToRunTimeSelectionTable(convectiveFluxScheme, roeB200FluxScheme, dictionary);
    addToRunTimeSelectionTable(convectiveFluxScheme, ausmPlusUpB200FluxScheme, dictionary);
    addToRunTimeSelectionTable(convectiveFluxScheme, rusanovB200FluxScheme, dictionary);
}

template<int Scheme>
Foam::B200FluxScheme<Scheme>::B200FluxScheme
(
    const dictionary& dict,
    const psiThermo& thermo,
    const volScalarField& rho,
    const volVectorField& U,
    const volScalarField& p
)
:
    convectiveFluxScheme(typeName, dict, thermo, rho, U, p),
    schemesSet_(false)
{}

template<int Scheme>
int Foam::B200FluxScheme<Scheme>::limiterId(const word& name)
{
    if (name == "vanLeer") return ICSB200_LIM_VANLEER;
    if (name == "Minmod") return ICSB200_LIM_MINMOD;
    if (name == "upwind") return ICSB200_LIM_UPWIND;
    if (name == "linear") return ICSB200_LIM_LINEAR;
    FatalErrorInFunction << "reconstruction scheme " << name << " is not available on the device (vanLeer Minmod upwind linear)" << exit(FatalError);
    return -1;
}

template<int Scheme>
void Foam::B200FluxScheme<Scheme>::setSchemes(const icsb200Mesh& dev) const
{
    if (schemesSet_) return;
    const fvMesh& m = mesh();
    const dictionary& interp = m.schemesDict().subDict("interpolationSchemes");
    const dictionary& pseudo = m.solutionDict().subDict("pseudoTime");
    const dictionary& cfs = dict().subDict("convectiveFluxScheme");
    icsb200_schemes s;
    s.flux_scheme = Scheme;
    s.limiter_rho = limiterId(word(interp.lookup("reconstruct(rho)")));
    s.limiter_U = limiterId(word(interp.lookup("reconstruct(U)")));
    s.limiter_T = limiterId(word(interp.lookup("reconstruct(T)")));
    s.low_mach_ausm = cfs.getOrDefault<Switch>("lowMachAusm", true);          // ausmPlusUpFluxScheme.C:61
    s.entropy_fix_coeff = cfs.getOrDefault<scalar>("entropyFixCoeff", 0.05);   // roeFluxScheme.C:255
    // ddtSchemes default: dualTime rPseudoDeltaT <inner>  (dualTimeDdtScheme.H:109-117)
    ITstream ddt(m.schemesDict().subDict("ddtSchemes").lookup("default"));
    const word inner(ddt.last().wordToken());
    s.ddt_scheme = inner == "steadyState" ? ICSB200_DDT_STEADY : inner == "Euler" ? ICSB200_DDT_EULER : ICSB200_DDT_BACKWARD;
    s.delta_t = m.time().deltaTValue();
    s.local_timestepping = pseudo.getOrDefault<Switch>("localTimestepping", true);
    s.local_timestepping_bounding = pseudo.getOrDefault<Switch>("localTimesteppingBounding", true);
    s.local_timestepping_lower_bound = min(max(pseudo.getOrDefault<scalar>("localTimesteppingLowerBound", 0.95), 0.0), 0.99);
    s.pseudo_co_num = pseudo.getOrDefault<scalar>("pseudoCoNum", 1.0);
    s.pseudo_co_num_min = pseudo.getOrDefault<scalar>("pseudoCoNumMin", 0.1);
    s.pseudo_co_num_max = pseudo.getOrDefault<scalar>("pseudoCoNumMax", 25.0);
    s.pseudo_co_num_max_incr = pseudo.getOrDefault<scalar>("pseudoCoNumMaxIncreaseFactor", 1.25);
    s.pseudo_co_num_min_decr = pseudo.getOrDefault<scalar>("pseudoCoNumMinDecreaseFactor", 0.1);
    s.rho_min = pseudo.getOrDefault<scalar>("rhoMin", -GREAT);
    s.T_min = pseudo.getOrDefault<scalar>("TMin", SMALL);
    s.T_max = pseudo.getOrDefault<scalar>("TMax", GREAT);
    s.viscous_full_jacobian = !dict().subOrEmptyDict("viscousFluxScheme").getOrDefault<Switch>("LaxFriedrichJacobian", true);
    dev.check(icsb200_schemes_set(dev.ctx(), &s), "icsb200_schemes_set");
    schemesSet_ = true;
}

template<int Scheme>
void Foam::B200FluxScheme<Scheme>::calcFlux(surfaceScalarField& phi, surfaceVectorField& phiUp, surfaceScalarField& phiEp)
{
    icsb200Mesh& dev = const_cast<icsb200Mesh&>(icsb200Mesh::New(mesh()));
    icsb200_ctx* c = dev.ctx();
    dev.setThermoAndBCs(thermo(), U());
    setSchemes(dev);

    // the solver has just assigned MRFFaceVelocity() / MRFOmega() (outerLoop.H:18-21): hand them down when a frame moves
    if (gMax(mag(MRFFaceVelocity().primitiveField())) > 0 || gMax(mag(MRFOmega().primitiveField())) > 0)
    {
        const scalarField fv(icsb200Mesh::flatten(MRFFaceVelocity()));
        dev.check(icsb200_mrf_set(c, fv.cdata(), &MRFOmega().primitiveField()[0].x()), "icsb200_mrf_set");
    }

    dev.uploadState(p(), U(), thermo().T());

    scalarField f(mesh().nFaces()), fe(mesh().nFaces());
    vectorField fu(mesh().nFaces());
    dev.check(icsb200_calc_flux(c, f.data(), &fu[0].x(), fe.data()), "icsb200_calc_flux");
    icsb200Mesh::unflatten(f, phi);
    icsb200Mesh::unflatten(fu, phiUp);
    icsb200Mesh::unflatten(fe, phiEp);

    // device-resident residual, pseudo time step and Jacobian for the same state: what residualsUpdate.H, setCoAndDeltaT.H and
    // createConvectiveJacobian compute on the host right after this call.  GMRESB200 uses them when its `deviceMatrix` switch is on.
    dev.check(icsb200_residual(c, nullptr, nullptr, nullptr), "icsb200_residual");
    dev.check(icsb200_pseudo_dt(c, nullptr, nullptr), "icsb200_pseudo_dt");
    dev.check(icsb200_assemble(c), "icsb200_assemble");
}

// explicit instantiation
template class Foam::B200FluxScheme<ICSB200_FLUX_HLLC>;
template class Foam::B200FluxScheme<ICSB200_FLUX_ROE>;
template class Foam::B200FluxScheme<ICSB200_FLUX_AUSMPLUSUP>;
