// icsb200Mesh.C — mesh, thermo and boundary-condition hand-off to libicsb200 (include/icsb200.h).
// NOT compiled in this repository's image (no OpenFOAM) — see README.md.
#include "icsb200Mesh.H"
#include "processorFvPatch.H"
#include "cyclicFvPatch.H"
#include "cyclicAMIFvPatch.H"
#include "emptyFvPatch.H"
#include "symmetryPlaneFvPatch.H"
#include "fixedValueFvPatchFields.H"
#include "zeroGradientFvPatchFields.H"
#include "slipFvPatchFields.H"
#include "symmetryPlaneFvPatchFields.H"
#include "inletOutletFvPatchFields.H"
#include "freestreamFvPatchFields.H"
#include "freestreamPressureFvPatchScalarField.H"
#include "totalPressureFvPatchScalarField.H"
#include "totalTemperatureFvPatchScalarField.H"
#include "pressureInletOutletVelocityFvPatchVectorField.H"
#include "Pstream.H"

namespace Foam
{
    defineTypeNameAndDebug(icsb200Mesh, 0);
}

void Foam::icsb200Mesh::check(const int rc, const char* where) const
{
    if (rc != 0)
    {
        FatalErrorInFunction
            << where << " failed (" << rc << "): " << (ctx_ ? icsb200_last_error(ctx_) : "no context")
            << exit(FatalError);
    }
}

Foam::icsb200Mesh::icsb200Mesh(const fvMesh& mesh)
:
    MeshObject<fvMesh, GeometricMeshObject, icsb200Mesh>(mesh),
    ctx_(nullptr),
    thermoSet_(false),
    stateIndex_(-1)
{
    // one context per MPI rank; the NCCL unique id is made on the master and scattered
    List<char> ncclId(128, '\0');
    if (Pstream::parRun())
    {
        if (Pstream::master()) icsb200_nccl_unique_id(ncclId.data());
        Pstream::scatter(ncclId);
    }
    // device of this rank: ranks of one node are numbered consecutively by mpirun; ICSB200_DEVICES_PER_NODE (default 8) GPUs per node
    const string perNodeEnv(Foam::getEnv("ICSB200_DEVICES_PER_NODE"));
    const label perNode = perNodeEnv.empty() ? 8 : readLabel(perNodeEnv);
    check
    (
        icsb200_create(&ctx_, Pstream::myProcNo() % perNode, Pstream::parRun() ? ncclId.cdata() : nullptr,
                       Pstream::myProcNo(), Pstream::nProcs()),
        "icsb200_create"
    );

    // ---- patches
    List<icsb200_patch> patches(mesh.boundary().size());
    forAll(mesh.boundary(), patchi)
    {
        const fvPatch& p = mesh.boundary()[patchi];
        icsb200_patch& q = patches[patchi];
        q.kind =
            isA<processorFvPatch>(p) ? ICSB200_PROCESSOR
          : isA<cyclicAMIFvPatch>(p) ? ICSB200_CYCLICAMI
          : isA<cyclicFvPatch>(p) ? ICSB200_CYCLIC
          : isA<emptyFvPatch>(p) ? ICSB200_EMPTY
          : (p.type() == "wall") ? ICSB200_WALL                    // setCoAndDeltaT.H:86 keys on the type NAME "wall"
          : isA<symmetryPlaneFvPatch>(p) ? ICSB200_SYMMETRYPLANE
          : ICSB200_PATCH;
        q.start = p.start();
        q.size = p.size();
        q.nbr_rank = -1;
        q.nbr_patch = -1;
        const tensor I(tensor::I);
        for (direction k = 0; k < 9; k++) q.forwardT[k] = I[k];
        if (isA<processorFvPatch>(p))
        {
            q.nbr_rank = refCast<const processorFvPatch>(p).neighbProcNo();
        }
        else if (isA<cyclicFvPatch>(p))
        {
            const cyclicFvPatch& cp = refCast<const cyclicFvPatch>(p);
            q.nbr_patch = cp.neighbFvPatch().index();
            if (cp.cyclicPatch().transform() == coupledPolyPatch::ROTATIONAL)
            {
                const tensor& T = cp.forwardT()[0];
                for (direction k = 0; k < 9; k++) q.forwardT[k] = T[k];
            }
        }
        else if (isA<cyclicAMIFvPatch>(p))
        {
            const cyclicAMIFvPatch& ap = refCast<const cyclicAMIFvPatch>(p);
            q.nbr_patch = ap.neighbFvPatch().index();
            if (ap.cyclicAMIPatch().transform() == coupledPolyPatch::ROTATIONAL)
            {
                const tensor& T = ap.forwardT()[0];
                for (direction k = 0; k < 9; k++) q.forwardT[k] = T[k];
            }
            // AMI addressing and weights of this side, flattened to CSR (cyclicAMIFvPatchField.C:146-209 interpolates with them)
            const AMIPatchToPatchInterpolation& ami = ap.owner() ? ap.AMI() : ap.neighbFvPatch().AMI();
            const labelListList& addr = ap.owner() ? ami.srcAddress() : ami.tgtAddress();
            const scalarListList& wght = ap.owner() ? ami.srcWeights() : ami.tgtWeights();
            labelList faceStart(p.size() + 1, 0);
            forAll(addr, i) faceStart[i + 1] = faceStart[i] + addr[i].size();
            labelList nbrFace(faceStart[p.size()]);
            scalarList weight(faceStart[p.size()]);
            forAll(addr, i) forAll(addr[i], k)
            {
                nbrFace[faceStart[i] + k] = addr[i][k];
                weight[faceStart[i] + k] = wght[i][k];
            }
            check(icsb200_ami_set(ctx_, patchi, p.size(), faceStart.cdata(), nbrFace.cdata(), weight.cdata()), "icsb200_ami_set");
        }
    }

    // ---- geometry: surface fields flattened internal-then-boundary; label is 32 bit, scalar is double (WM_LABEL_SIZE=32, DP)
    const vectorField Sf(flatten(mesh.Sf()));
    const scalarField magSf(flatten(mesh.magSf()));
    const scalarField w(flatten(mesh.weights()));
    const scalarField dc(flatten(mesh.deltaCoeffs()));
    const scalarField nodc(flatten(mesh.nonOrthDeltaCoeffs()));
    const vectorField Cf(flatten(mesh.Cf()));
    const Vector<label>& solD = mesh.solutionD();
    const int solutionD[3] = {int(solD.x()), int(solD.y()), int(solD.z())};
    check
    (
        icsb200_mesh_set
        (
            ctx_, mesh.nCells(), mesh.nInternalFaces(), mesh.nFaces(),
            mesh.faceOwner().cdata(), mesh.faceNeighbour().cdata(),
            &Sf[0].x(), magSf.cdata(), w.cdata(), dc.cdata(), nodc.cdata(),
            &mesh.C().primitiveField()[0].x(), mesh.V().field().cdata(), &Cf[0].x(),
            patches.size(), patches.cdata(), solutionD
        ),
        "icsb200_mesh_set"
    );
}

Foam::icsb200Mesh::~icsb200Mesh()
{
    if (ctx_) icsb200_destroy(ctx_);
}

bool Foam::icsb200Mesh::movePoints()
{
    FatalErrorInFunction << "moving meshes are not supported by libicsb200 (DESIGN.md §7)" << exit(FatalError);
    return false;
}

void Foam::icsb200Mesh::setThermoAndBCs(const psiThermo& thermo, const volVectorField& U)
{
    if (thermoSet_) return;
    const fvMesh& mesh = U.mesh();
    // constant-property perfect gas: read back from the thermo fields of cell 0 (hConst / constTransport / perfectGas)
    const scalar Cp = thermo.Cp()().primitiveField()[0];
    const scalar Cv = thermo.Cv()().primitiveField()[0];
    const scalar mu = thermo.mu()().primitiveField()[0];
    const scalar alpha = thermo.alpha().primitiveField()[0];                 // kappa / Cp
    check(icsb200_thermo_set(ctx_, Cp - Cv, Cp, mu, alpha > VSMALL ? mu/alpha : 1.0), "icsb200_thermo_set");

    const volScalarField& p = thermo.p();
    const volScalarField& T = thermo.T();
    const scalar gamma = Cp/Cv;
    auto uniformOr = [&](const label patchi, const int field, const int kind, const scalarField& rows, const label nPrm)
    {
        // rows: nPrm parameters per face; one bc_set when all faces agree, the per-face variant otherwise
        const label n = rows.size()/max(nPrm, label(1));
        bool uni = true;
        for (label i = 1; i < n && uni; i++) for (label k = 0; k < nPrm; k++) uni = uni && rows[i*nPrm + k] == rows[k];
        if (nPrm == 0 || n == 0) check(icsb200_bc_set(ctx_, patchi, field, kind, nullptr, 0), "icsb200_bc_set");
        else if (uni) check(icsb200_bc_set(ctx_, patchi, field, kind, rows.cdata(), nPrm), "icsb200_bc_set");
        else check(icsb200_bc_set_nonuniform(ctx_, patchi, field, kind, rows.cdata(), nPrm), "icsb200_bc_set_nonuniform");
    };
    auto rows3 = [](const vectorField& v)
    {
        scalarField r(3*v.size());
        forAll(v, i) for (direction d = 0; d < 3; d++) r[3*i + d] = v[i][d];
        return r;
    };
    forAll(mesh.boundary(), patchi)
    {
        const fvPatch& fp = mesh.boundary()[patchi];
        if (fp.coupled() || isA<emptyFvPatch>(fp)) continue;                 // set automatically by the library
        const label n = fp.size();
        // ---- p
        {
            const fvPatchScalarField& pf = p.boundaryField()[patchi];
            if (isA<totalPressureFvPatchScalarField>(pf))
            {
                const scalarField& p0 = refCast<const totalPressureFvPatchScalarField>(pf).p0();
                scalarField rows(2*n);
                forAll(p0, i) { rows[2*i] = p0[i]; rows[2*i + 1] = gamma; }
                uniformOr(patchi, ICSB200_FIELD_P, ICSB200_BC_TOTALPRESSURE, rows, 2);
            }
            else if (isA<freestreamPressureFvPatchScalarField>(pf))
            {
                const freestreamPressureFvPatchScalarField& fpf = refCast<const freestreamPressureFvPatchScalarField>(pf);
                const freestreamFvPatchVectorField& Uf = refCast<const freestreamFvPatchVectorField>(U.boundaryField()[patchi]);
                scalarField rows(4*n);
                forAll(fpf, i)
                {
                    rows[4*i] = fpf.freestreamValue()[i];
                    for (direction d = 0; d < 3; d++) rows[4*i + 1 + d] = Uf.freestreamValue()[i][d];
                }
                uniformOr(patchi, ICSB200_FIELD_P, ICSB200_BC_FREESTREAMPRESSURE, rows, 4);
            }
            else if (isA<fixedValueFvPatchScalarField>(pf)) uniformOr(patchi, ICSB200_FIELD_P, ICSB200_BC_FIXEDVALUE, scalarField(pf), 1);
            else if (isA<zeroGradientFvPatchScalarField>(pf) || isA<symmetryPlaneFvPatchScalarField>(pf) || isA<slipFvPatchScalarField>(pf))
                uniformOr(patchi, ICSB200_FIELD_P, ICSB200_BC_ZEROGRADIENT, scalarField(), 0);
            else FatalErrorInFunction << "patch field type " << pf.type() << " of p on " << fp.name() << " is not supported" << exit(FatalError);
        }
        // ---- U
        {
            const fvPatchVectorField& uf = U.boundaryField()[patchi];
            if (isA<pressureInletOutletVelocityFvPatchVectorField>(uf))
            {
                const pressureInletOutletVelocityFvPatchVectorField& pu = refCast<const pressureInletOutletVelocityFvPatchVectorField>(uf);
                // refValue = tangentialVelocity - n (n & tangentialVelocity) is rebuilt on the device from the raw entry; a case
                // without the entry hands over zeros
                uniformOr(patchi, ICSB200_FIELD_U, ICSB200_BC_PRESSUREINLETOUTLETVELOCITY, rows3(pu.refValue()), 3);
            }
            else if (isA<freestreamFvPatchVectorField>(uf))
                uniformOr(patchi, ICSB200_FIELD_U, ICSB200_BC_INLETOUTLET, rows3(refCast<const freestreamFvPatchVectorField>(uf).freestreamValue()), 3);
            else if (isA<inletOutletFvPatchVectorField>(uf))
                uniformOr(patchi, ICSB200_FIELD_U, ICSB200_BC_INLETOUTLET, rows3(refCast<const inletOutletFvPatchVectorField>(uf).refValue()), 3);
            else if (isA<slipFvPatchVectorField>(uf) || isA<symmetryPlaneFvPatchVectorField>(uf))
                uniformOr(patchi, ICSB200_FIELD_U, ICSB200_BC_SLIP, scalarField(), 0);
            else if (isA<fixedValueFvPatchVectorField>(uf)) uniformOr(patchi, ICSB200_FIELD_U, ICSB200_BC_FIXEDVALUE, rows3(uf), 3);    // noSlip is a fixedValue
            else if (isA<zeroGradientFvPatchVectorField>(uf)) uniformOr(patchi, ICSB200_FIELD_U, ICSB200_BC_ZEROGRADIENT, scalarField(), 0);
            else FatalErrorInFunction << "patch field type " << uf.type() << " of U on " << fp.name() << " is not supported" << exit(FatalError);
        }
        // ---- T
        {
            const fvPatchScalarField& tf = T.boundaryField()[patchi];
            if (isA<totalTemperatureFvPatchScalarField>(tf))
            {
                const scalarField& T0 = refCast<const totalTemperatureFvPatchScalarField>(tf).T0();
                scalarField rows(2*n);
                forAll(T0, i) { rows[2*i] = T0[i]; rows[2*i + 1] = gamma; }
                uniformOr(patchi, ICSB200_FIELD_T, ICSB200_BC_TOTALTEMPERATURE, rows, 2);
            }
            else if (isA<inletOutletFvPatchScalarField>(tf))
                uniformOr(patchi, ICSB200_FIELD_T, ICSB200_BC_INLETOUTLET, scalarField(refCast<const inletOutletFvPatchScalarField>(tf).refValue()), 1);
            else if (isA<fixedValueFvPatchScalarField>(tf)) uniformOr(patchi, ICSB200_FIELD_T, ICSB200_BC_FIXEDVALUE, scalarField(tf), 1);
            else if (isA<zeroGradientFvPatchScalarField>(tf) || isA<symmetryPlaneFvPatchScalarField>(tf) || isA<slipFvPatchScalarField>(tf))
                uniformOr(patchi, ICSB200_FIELD_T, ICSB200_BC_ZEROGRADIENT, scalarField(), 0);
            else FatalErrorInFunction << "patch field type " << tf.type() << " of T on " << fp.name() << " is not supported" << exit(FatalError);
        }
    }
    thermoSet_ = true;
}

void Foam::icsb200Mesh::uploadState(const volScalarField& p, const volVectorField& U, const volScalarField& T)
{
    check(icsb200_state_set(ctx_, p.primitiveField().cdata(), &U.primitiveField()[0].x(), T.primitiveField().cdata()), "icsb200_state_set");
}
