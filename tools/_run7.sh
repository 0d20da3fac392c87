cd oracle/_ref/reftests
mkdir -p ../../../gpurun_out/reftests
for f in test_mappings test_lagrange test_dof_ordering test_lattice test_equality test_polynomials test_create test_entity_dofs test_interpolation_between_elements test_orthonormal_basis test_tensor_products test_nedelec test_rt test_dof_transformations; do
  s=$(date +%s)
  timeout 900 python -m pytest test/$f.py -q -p no:cacheprovider -rf > ../../../gpurun_out/reftests/$f.log 2>&1
  echo "$(tail -1 ../../../gpurun_out/reftests/$f.log) [$(( $(date +%s) - s )) s] <- $f"
done
