"""Structural check of the Fortran iso_c_binding shim (fortran/mlf_emgmm_b200.f90).  No Fortran compiler exists in
this image, so the shim cannot be compiled here; what can be checked is that every bind(C) name it declares is
exported by libmlfortran_b200.so and that the number of dummy arguments equals the number of parameters of the C
prototype in include/mlf_b200.h."""
import os
import re

from mlfortran_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_prototypes():
    src = open(os.path.join(ROOT, "include", "mlf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(mlf_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_bind_c_names_match_the_c_abi():
    f90 = open(os.path.join(ROOT, "fortran", "mlf_emgmm_b200.f90")).read()
    f90 = re.sub(r"&\s*\n\s*", " ", f90)                      # join continuation lines
    binds = re.findall(r"(?:Function|Subroutine)\s+\w+\s*\(([^)]*)\)[^\n]*bind\(C,\s*name=\"(\w+)\"\)", f90, flags=re.I)
    assert len(binds) >= 6
    protos = _c_prototypes()
    lib = _lib.load()
    for args, name in binds:
        assert hasattr(lib, name), f"{name} is bound by the Fortran shim but not exported"
        nargs = len([a for a in args.split(",") if a.strip()])
        assert protos[name] == nargs, f"{name}: Fortran passes {nargs} arguments, the C prototype takes {protos[name]}"


def test_shim_mirrors_the_reference_type_interface():
    f90 = open(os.path.join(ROOT, "fortran", "mlf_emgmm_b200.f90")).read()
    # public components and type-bound procedures a caller of mlf_algo_emgmm relies on (src/unsupervised/mlf_emgmm.f90:46-56)
    for token in ("Extends(mlf_step_obj)", "Mu(:,:)", "Cov(:,:,:)", "lambda(:)", "sumLL", "minProba", "initialized",
                  "procedure :: init =>", "procedure :: reinit =>", "procedure :: stepF =>", "getProba =>", "getNumClasses =>"):
        assert token in f90, token
    # same slot names and restore rule as the reference init (:90-101)
    for token in ('C_CHAR_"Mu"', 'C_CHAR_"Cov"', 'C_CHAR_"lambda"', '"sumLL"', '"minProba"', "fixed_dims = [.TRUE., .FALSE.]"):
        assert token in f90, token


def test_example_program_uses_only_what_the_shim_exports():
    """fortran/example_emgmm_b200.f90 (the reference's testGMM switched to the library) calls only public names of the shim
    module and type-bound procedures / components the shim type has."""
    f90 = open(os.path.join(ROOT, "fortran", "mlf_emgmm_b200.f90")).read()
    ex = open(os.path.join(ROOT, "fortran", "example_emgmm_b200.f90")).read()
    assert "Use mlf_emgmm_b200" in ex and "type(mlf_algo_emgmm_b200) :: em" in ex
    for name in ("mlf_b200_synth_gmm_f", "mlf_b200_use_all_gpus"):
        assert name in ex and re.search(r"Public\s*(::)?[^\n]*\b" + name + r"\b", f90, flags=re.I), name
    for member in re.findall(r"\bem%(\w+)", ex):
        # own components / bindings, or inherited from the reference's mlf_step_obj (step) and mlf_obj (finalize)
        assert re.search(r"\b" + member + r"\b", f90) or member in ("step",), member
