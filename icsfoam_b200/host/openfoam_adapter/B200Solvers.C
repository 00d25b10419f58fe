// NOT compiled in this repository's image (no OpenFOAM) — see README.md.
#include "B200Solvers.H"
#include "addToRunTimeSelectionTable.H"

namespace Foam
{
    defineTemplateTypeNameAndDebugWithName(gmresB200, "GMRESB200", 0);
    defineTemplateTypeNameAndDebugWithName(smoothSolverCoupledB200, "smoothSolverCoupledB200", 0);
    coupledMatrix::solver::adddictionaryConstructorToTable<gmresB200> addgmresB200DictionaryConstructorToTable_;
    coupledMatrix::solver::adddictionaryConstructorToTable<smoothSolverCoupledB200> addsmoothB200DictionaryConstructorToTable_;

    defineTemplateTypeNameAndDebugWithName(lusgsB200, "LUSGSB200", 0);
    defineTemplateTypeNameAndDebugWithName(jacobiB200, "JacobiB200", 0);
    coupledMatrix::preconditioner::adddictionaryConstructorToTable<lusgsB200> addlusgsB200DictionaryConstructorToTable_;
    coupledMatrix::preconditioner::adddictionaryConstructorToTable<jacobiB200> addjacobiB200DictionaryConstructorToTable_;

    //- the nine LDU sub-blocks in the numbering of icsb200_matrix_set_ldu (coupledMatrix.H:399-421), and the flattened
    //  interfacesUpper of the coupled patches (blockFvMatrix.C:248-266)
    template<class sT, class bT>              // bT: scalar, vector or tensor coefficient (blockFvMatrix.H:57-73)
    static void pushBlock(const icsb200Mesh& dev, const int id, const blockFvMatrix<sT, bT>& b)
    {
        const double* d = b.hasDiag() ? reinterpret_cast<const double*>(b.diag().cdata()) : nullptr;
        const double* u = b.hasUpper() ? reinterpret_cast<const double*>(b.upper().cdata()) : nullptr;
        const double* l = b.hasLower() ? reinterpret_cast<const double*>(b.lower().cdata()) : nullptr;
        dev.check(icsb200_matrix_set_ldu(dev.ctx(), id, d, u, l), "icsb200_matrix_set_ldu");
        const fvMesh& mesh = dev.mesh();
        if (b.interfacesUpper().size())
        {
            Field<bT> flat(mesh.nFaces() - mesh.nInternalFaces(), Zero);
            forAll(mesh.boundary(), patchi)
            {
                const fvPatch& p = mesh.boundary()[patchi];
                if (p.coupled() && b.interfacesUpper().set(patchi))
                {
                    SubList<bT>(flat, p.size(), p.start() - mesh.nInternalFaces()) = b.interfacesUpper()[patchi];
                }
            }
            dev.check(icsb200_matrix_set_interfaces(dev.ctx(), id, reinterpret_cast<const double*>(flat.cdata())), "icsb200_matrix_set_interfaces");
        }
    }
}

template<int Solver>
Foam::B200Solver<Solver>::B200Solver(const dictionary& dict, const coupledMatrix& matrix)
:
    coupledMatrix::solver(typeName, dict, matrix),
    nDirsOrSweeps_(Solver == ICSB200_SOLVER_GMRES ? controlDict_.get<label>("nDirections") : controlDict_.getOrDefault<label>("nSweeps", 1)),
    preconditioner_(ICSB200_PRECOND_JACOBI),
    deviceMatrix_(controlDict_.getOrDefault<Switch>("deviceMatrix", false))
{
    if (Solver == ICSB200_SOLVER_GMRES)
    {
        const word pc(controlDict_.get<word>("preconditioner"));
        if (pc == "LUSGS" || pc == "LUSGSB200") preconditioner_ = ICSB200_PRECOND_LUSGS;
        else if (pc == "Jacobi" || pc == "JacobiB200") preconditioner_ = ICSB200_PRECOND_JACOBI;
        else FatalErrorInFunction << "Unknown preconditioner " << pc << nl << "Valid preconditioners are : LUSGS Jacobi" << exit(FatalError);
    }
}

template<int Solver>
void Foam::B200Solver<Solver>::uploadMatrix(const icsb200Mesh& dev) const
{
    const coupledMatrix& A = matrix();
    pushBlock(dev, 0, A.dSByS(0,0));
    pushBlock(dev, 1, A.dSByS(0,1));
    pushBlock(dev, 2, A.dSByS(1,0));
    pushBlock(dev, 3, A.dSByS(1,1));
    pushBlock(dev, 4, A.dSByV(0,0));
    pushBlock(dev, 5, A.dSByV(1,0));
    pushBlock(dev, 6, A.dVByS(0,0));
    pushBlock(dev, 7, A.dVByS(0,1));
    pushBlock(dev, 8, A.dVByV(0,0));
}

template<int Solver>
Foam::residualsIO Foam::B200Solver<Solver>::solveDelta
(
    PtrList<volScalarField>& sW, PtrList<volVectorField>& vW,
    const PtrList<scalarField>& sSource, const PtrList<vectorField>& vSource,
    PtrList<volScalarField>& dsW, PtrList<volVectorField>& dvW
) const
{
    const icsb200Mesh& dev = icsb200Mesh::New(matrix().mesh());
    icsb200_ctx* c = dev.ctx();
    if (!deviceMatrix_)
    {
        uploadMatrix(dev);
    }
    // the sources carry everything the solver added on the host (fvOptions, HB, ...): always taken from the caller
    dev.check(icsb200_source_set(c, sSource[0].cdata(), &vSource[0][0].x(), sSource[1].cdata()), "icsb200_source_set");

    icsb200_solver_controls ctl;
    ctl.solver = Solver;
    ctl.preconditioner = preconditioner_;
    ctl.n_directions = nDirsOrSweeps_;
    ctl.max_iter = maxIter_;
    ctl.min_iter = minIter_;
    ctl.tolerance = tolerance_;
    ctl.rel_tol = relTolerance_;
    icsb200_residuals r;
    dev.check
    (
        icsb200_solve_delta(c, &ctl, dsW[0].primitiveFieldRef().data(), &dvW[0].primitiveFieldRef()[0].x(), dsW[1].primitiveFieldRef().data(), &r),
        "icsb200_solve_delta"
    );
    residualsIO perf(2, 1);                                        // residualsIO.H:74
    for (label i = 0; i < 2; i++)
    {
        perf.sInitRes()[i] = r.s_init[i];
        perf.sFinalRes()[i] = r.s_final[i];
    }
    perf.vInitRes()[0] = vector(r.v_init[0], r.v_init[1], r.v_init[2]);
    perf.vFinalRes()[0] = vector(r.v_final[0], r.v_final[1], r.v_final[2]);
    perf.nIterations() = r.n_iterations;
    return perf;
}

template<int Solver>
Foam::residualsIO Foam::B200Solver<Solver>::solveDelta
(
    PtrList<volScalarField>& sW, PtrList<volVectorField>& vW,
    const PtrList<scalarField>& sSource, const PtrList<vectorField>& vSource
) const
{
    // as gmres.C:260-330: increments in temporaries, then W += dW
    PtrList<volScalarField> dsW(sW.size());
    PtrList<volVectorField> dvW(vW.size());
    forAll(sW, i) dsW.set(i, new volScalarField("d" + sW[i].name(), sW[i]));
    forAll(vW, i) dvW.set(i, new volVectorField("d" + vW[i].name(), vW[i]));
    residualsIO perf(solveDelta(sW, vW, sSource, vSource, dsW, dvW));
    forAll(sW, i) sW[i].primitiveFieldRef() += dsW[i].primitiveField();
    forAll(vW, i) vW[i].primitiveFieldRef() += dvW[i].primitiveField();
    return perf;
}

template<int Solver>
Foam::residualsIO Foam::B200Solver<Solver>::solve
(
    PtrList<volScalarField>&, PtrList<volVectorField>&,
    const PtrList<scalarField>&, const PtrList<vectorField>&
) const
{
    // the dbnsFoam path is in delta form (coupledMatrix(mesh, 2, 1, true), outerLoop.H:53) and never calls solve()
    FatalErrorInFunction << "only the delta form (solveDelta) is available on the device" << exit(FatalError);
    return residualsIO(2, 1);
}

template<int Precond>
Foam::B200Preconditioner<Precond>::B200Preconditioner(const coupledMatrix::solver& sol, const dictionary&)
:
    coupledMatrix::preconditioner(sol),
    matrix_(sol.matrix())
{
    // a host solver (the reference's GMRES) applies this preconditioner: the device needs the host matrix once per solve
    const icsb200Mesh& dev = icsb200Mesh::New(matrix_.mesh());
    pushBlock(dev, 0, matrix_.dSByS(0,0)); pushBlock(dev, 1, matrix_.dSByS(0,1)); pushBlock(dev, 2, matrix_.dSByS(1,0));
    pushBlock(dev, 3, matrix_.dSByS(1,1)); pushBlock(dev, 4, matrix_.dSByV(0,0)); pushBlock(dev, 5, matrix_.dSByV(1,0));
    pushBlock(dev, 6, matrix_.dVByS(0,0)); pushBlock(dev, 7, matrix_.dVByS(0,1)); pushBlock(dev, 8, matrix_.dVByV(0,0));
}

template<int Precond>
void Foam::B200Preconditioner<Precond>::precondition(PtrList<scalarField>& sVec, PtrList<vectorField>& vVec) const
{
    const icsb200Mesh& dev = icsb200Mesh::New(matrix_.mesh());
    dev.check(icsb200_precondition(dev.ctx(), Precond, sVec[0].data(), &vVec[0][0].x(), sVec[1].data()), "icsb200_precondition");
}

template class Foam::B200Solver<ICSB200_SOLVER_GMRES>;
template class Foam::B200Solver<ICSB200_SOLVER_SMOOTH>;
template class Foam::B200Preconditioner<ICSB200_PRECOND_LUSGS>;
template class Foam::B200Preconditioner<ICSB200_PRECOND_JACOBI>;
