// Intentionally empty: shadows the reference header of the same name, which is not on the
// leaf-sweep path and would pull in unrelated dune-common headers.
