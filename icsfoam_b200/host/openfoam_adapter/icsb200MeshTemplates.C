// template members of icsb200Mesh (flattening of surface / boundary fields into the face order of icsb200_mesh_set)
template<class Type>
Foam::Field<Type> Foam::icsb200Mesh::flatten(const GeometricField<Type, fvsPatchField, surfaceMesh>& sf)
{
    const fvMesh& mesh = sf.mesh();
    Field<Type> flat(mesh.nFaces(), Zero);
    SubList<Type>(flat, mesh.nInternalFaces()) = sf.primitiveField();
    forAll(mesh.boundary(), patchi)
    {
        const fvPatch& p = mesh.boundary()[patchi];
        if (sf.boundaryField()[patchi].size() == p.size())       // empty patches carry zero-size fields
        {
            SubList<Type>(flat, p.size(), p.start()) = sf.boundaryField()[patchi];
        }
    }
    return flat;
}

template<class Type>
void Foam::icsb200Mesh::unflatten(const Field<Type>& flat, GeometricField<Type, fvsPatchField, surfaceMesh>& sf)
{
    const fvMesh& mesh = sf.mesh();
    sf.primitiveFieldRef() = SubList<Type>(flat, mesh.nInternalFaces());
    forAll(mesh.boundary(), patchi)
    {
        const fvPatch& p = mesh.boundary()[patchi];
        if (sf.boundaryField()[patchi].size() == p.size())
        {
            sf.boundaryFieldRef()[patchi] = SubList<Type>(flat, p.size(), p.start());
        }
    }
}

template<class Type>
Foam::Field<Type> Foam::icsb200Mesh::flattenBoundary(const GeometricField<Type, fvPatchField, volMesh>& vf)
{
    const fvMesh& mesh = vf.mesh();
    Field<Type> flat(mesh.nFaces() - mesh.nInternalFaces(), Zero);
    forAll(mesh.boundary(), patchi)
    {
        const fvPatch& p = mesh.boundary()[patchi];
        if (vf.boundaryField()[patchi].size() == p.size())
        {
            SubList<Type>(flat, p.size(), p.start() - mesh.nInternalFaces()) = vf.boundaryField()[patchi];
        }
    }
    return flat;
}
