"""Unit tests of the Fortran -> C translator (oracle/f90_to_c.py) on a small Fortran module WRITTEN FOR THIS TEST (no model code):
the Fortran semantics the pin depends on -- operator precedence and associativity, default-real literals, integer division,
whole-array statements, where / elsewhere, elemental calls over arrays, sum() over sections, optional arguments, blank-padded
character compares, select case -- checked against values computed independently in Python.  CPU only; needs gcc."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")

SRC = """
module probe
    implicit none
    integer, parameter :: dp = kind(0.d0)
    double precision, parameter :: two = 2.0_dp
    type box_class
        double precision, allocatable, dimension(:) :: a, b
        integer, allocatable, dimension(:) :: m
        character(len=16) :: tag
        double precision :: s
    end type
contains
    ! scalar expression semantics
    subroutine exprs(x, y, n, out)
        double precision, intent(in) :: x, y
        integer, intent(in) :: n
        double precision, intent(out) :: out(12)
        out(1) = -x**2                 ! -(x**2)
        out(2) = 2**3**2               ! right associative: 2**9
        out(3) = x/y*two               ! (x/y)*2
        out(4) = x - y - two           ! (x-y)-2
        out(5) = -x*y + two            ! (-(x*y)) + 2
        out(6) = 0.8                   ! default real literal promoted to double
        out(7) = n/2*2                 ! integer division
        out(8) = x**2.0_dp + y**3      ! real and integer exponents
        out(9) = dmax1(0.0_dp, -x) + dmin1(x, y)
        out(10) = abs(-y)*dble(n)/3
        out(11) = (x + y)*(x - y)
        out(12) = mod(n, 5) + n**2
    end subroutine

    elemental subroutine clip(v, lo, hi, r)
        double precision, intent(in) :: v, lo, hi
        double precision, intent(out) :: r
        r = v
        if (v < lo) r = lo
        if (v > hi) then
            r = hi
        end if
    end subroutine

    elemental function cube(v) result(r)
        double precision, intent(in) :: v
        double precision :: r
        r = v*v*v
    end function

    ! whole-array statements, where/elsewhere, elemental calls, character logic
    subroutine arrays(box, n, scale, nhit)
        type(box_class), intent(in out) :: box
        integer, intent(in) :: n
        double precision, optional, intent(in) :: scale
        integer, intent(out) :: nhit
        double precision, dimension(n) :: tmp
        integer :: q
        tmp = box%a*two - 1.0_dp
        call clip(tmp, -0.5_dp, 0.5_dp, box%b)
        where (box%m == 2 .or. box%a > 0.75_dp)
            box%b = box%b + cube(box%a)
            box%a = 0.0_dp
        elsewhere
            box%b = -box%b
        end where
        if (present(scale)) then
            box%b = box%b*scale
        end if
        box%s = sum(box%b(:))/n
        nhit = 0
        do q = 1, n
            select case (box%m(q))
                case (2)
                    nhit = nhit + 10
                case (0, 1)
                    nhit = nhit + 1
                case default
                    nhit = nhit - 100
            end select
        end do
        if (trim(box%tag) .eq. "ice") nhit = nhit + 1000
        if (box%tag .eq. "ice   ") nhit = nhit + 2000          ! blank-padded compare
        if (trim(box%tag) .ne. "icy") nhit = nhit + 4000
    end subroutine

    function colmean(x, nx, nt) result(r)
        integer, intent(in) :: nx, nt
        double precision, intent(in), dimension(nx,nt) :: x
        double precision :: r(nx)
        integer i
        do i=1,nx
            r(i) = sum((x(i,:) - x(i,1))**2)/nt
        end do
    end function

    subroutine box_alloc(box, n)
        type(box_class) :: box
        integer :: n
        allocate(box%a(n))
        allocate(box%b(n))
        allocate(box%m(n))
    end subroutine
end module probe
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("f90")
    f90, c, so = d / "probe.f90", d / "probe.c", d / "libprobe.so"
    f90.write_text(SRC)
    subprocess.check_call([sys.executable, os.path.join(ORACLE, "f90_to_c.py"), "--prefix", "p_", "-o", str(c), str(f90)])
    subprocess.check_call(["gcc", "-O2", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I", ORACLE, "-o", str(so),
                           str(c), os.path.join(ORACLE, "f90_rt.c"), os.path.join(ORACLE, "ref_harness.c"), "-lm",
                           "-Wl,--defsym=ref_types=p_types,--defsym=ref_surface_energy_and_mass_balance=p_exprs"])
    return C.CDLL(str(so)), c.read_text()


def test_expression_semantics(lib):
    L, _ = lib
    dp = C.POINTER(C.c_double)
    out = np.zeros(12)
    x, y, n = 1.7, -0.6, 7
    L.p_exprs(C.byref(C.c_double(x)), C.byref(C.c_double(y)), C.byref(C.c_int(n)), out.ctypes.data_as(dp))
    want = [-(x * x), 512.0, (x / y) * 2.0, (x - y) - 2.0, -(x * y) + 2.0, float(np.float32(0.8)), (n // 2) * 2,
            x * x + (y * y) * y, max(0.0, -x) + min(x, y), abs(-y) * float(n) / 3, (x + y) * (x - y), n % 5 + n * n]
    assert out.tolist() == want


class Arr(C.Structure):
    _fields_ = [("a", C.c_void_p), ("n0", C.c_long), ("n1", C.c_long)]


def test_array_statements_where_elemental_optional_and_characters(lib):
    L, text = lib
    n = 9
    L.ref_new.restype = C.c_void_p
    L.ref_new.argtypes = [C.c_char_p]
    L.ref_comp.restype = C.c_void_p
    L.ref_comp.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
    box = L.ref_new(b"box_class")
    L.p_box_alloc.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.p_box_alloc(box, C.byref(C.c_int(n)))

    def view(name, ctype, dtype):
        d = Arr.from_address(L.ref_comp(b"box_class", box, name, None, None))
        return np.ctypeslib.as_array(C.cast(d.a, C.POINTER(ctype)), shape=(d.n0,))

    a, b, m = view(b"a", C.c_double, np.float64), view(b"b", C.c_double, np.float64), view(b"m", C.c_int, np.int32)
    r = np.random.default_rng(0)
    a[:] = r.uniform(0, 1, n)
    m[:] = [2, 0, 1, 3, 2, 1, 0, 2, 5]
    C.memmove(L.ref_comp(b"box_class", box, b"tag", None, None), b"ice".ljust(16), 16)
    a0 = a.copy()
    nhit = C.c_int()
    L.p_arrays.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.p_arrays(box, C.byref(C.c_int(n)), C.byref(C.c_double(3.0)), C.byref(nhit))
    tmp = a0 * 2.0 - 1.0
    want_b = np.clip(tmp, -0.5, 0.5)
    mask = (m == 2) | (a0 > 0.75)                      # evaluated BEFORE the body changes box%a
    want_b = np.where(mask, want_b + a0 * a0 * a0, -want_b) * 3.0
    assert np.array_equal(b, want_b)
    assert np.array_equal(a, np.where(mask, 0.0, a0))
    s = 0.0
    for v in want_b:
        s = s + v                                       # sequential, like gfortran's sum without -ffast-math
    assert C.c_double.from_address(L.ref_comp(b"box_class", box, b"s", None, None)).value == s / n
    assert nhit.value == 3 * 10 + 4 * 1 - 2 * 100 + 1000 + 2000 + 4000
    # optional argument absent: no scaling
    a[:] = a0
    L.p_arrays(box, C.byref(C.c_int(n)), None, C.byref(nhit))
    assert np.array_equal(b, want_b / 3.0)
    # the mask of a where construct lives in a temporary, one loop per body statement
    assert "unsigned char *m" in text and text.count("if (m") + text.count("if (!m") >= 3


def test_array_function_with_sections(lib):
    L, _ = lib
    nx, nt = 4, 6
    x = np.asfortranarray(np.random.default_rng(1).normal(size=(nx, nt)))
    r = np.zeros(nx)
    dp = C.POINTER(C.c_double)
    L.p_colmean(x.ctypes.data_as(dp), C.byref(C.c_int(nx)), C.byref(C.c_int(nt)), r.ctypes.data_as(dp))
    for i in range(nx):
        s = 0.0
        for j in range(nt):
            d = x[i, j] - x[i, 0]
            s = s + d * d
        assert r[i] == s / nt


def test_generated_code_cites_lines_and_keeps_no_source_text(lib):
    _, text = lib
    assert "probe.f90:" in text and "blank-padded compare" not in text and "right associative" not in text
