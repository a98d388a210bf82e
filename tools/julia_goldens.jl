# Digit-level golden vectors from the REAL package (run by anyone who has Julia; this image has none).
#
#   julia --project=<env with CurvilinearGrids v0.13.3> tools/julia_goldens.jl [repo_root]
#
# Reads the cases of tests/golden/julia_inputs/<case>/ (written by tools/make_julia_inputs.py: raw little-endian
# Float64 arrays in Julia memory order + case.txt), builds every grid with the STOCK public API (DiscreteGrid /
# MappedGrid constructors, cell_metrics, face_metrics, node / centroid / face coordinates) and writes, per case and
# conserved-metric scheme, tests/golden/julia/<case>/<scheme>/{cell_forward,cell_inverse,face<a>_forward,
# face<a>_inverse,face<a>_conserved,node<c>,centroid<c>,face<a>coord<c>}.bin — `reinterpret(Float64, vec(array))`,
# i.e. exactly the memory libcurvigrid_cuda.so writes (include/curvigrid_cuda.h: element = N*N matrix entries
# column-major, then J).  tests/test_julia_goldens.py then checks the oracle AND the CUDA path against these digits;
# until this has been run the oracle is pinned by the reference's re-encoded property tests only ("parity unpinned").
using CurvilinearGrids

const ROOT = length(ARGS) >= 1 ? ARGS[1] : normpath(joinpath(@__DIR__, ".."))
const INP = joinpath(ROOT, "tests", "golden", "julia_inputs")
const OUT = joinpath(ROOT, "tests", "golden", "julia")

function read_case(dir)
  meta = Dict{String,String}()
  for ln in eachline(joinpath(dir, "case.txt"))
    isempty(strip(ln)) && continue
    k, v = split(ln, "="; limit=2)
    meta[strip(k)] = strip(v)
  end
  return meta
end
ints(s) = Tuple(parse.(Int, split(s, ",")))
readf64(path, dims...) = reshape(collect(reinterpret(Float64, read(path))), dims...)

scheme_of(s) = s == "endpoint" ? EndpointAverageReconstruction() :
               s == "gradient" ? GradientCorrectedReconstruction() :
               s == "curvature" ? CurvatureCorrectedReconstruction() :
               s == "adtl" ? ADThomasLombardMetric() : error("unknown scheme $s")
cs_of(s) = s == "SphericalCS" ? SphericalCS() : s == "CylindricalCS" ? CylindricalCS() : CurvilinearCS()
basis_of(s) = s == "SphericalBasis" ? SphericalBasis() : CartesianBasis()

# separable term list (include/curvigrid_cuda.h, cg_grid_set_mapping_terms): 15 x nterms matrix, per term
# [coord, c0, c1, c2, w, ph, (kind, a, b) x 3]; kind 0: 1, 1: a + b s, 2: sin(a + b s), 3: cos(a + b s)
factor(kind, a, b, s) = kind == 0 ? one(s) : kind == 1 ? a + b * s : kind == 2 ? sin(a + b * s) : cos(a + b * s)
function term_mapping(terms::Matrix{Float64}, coord::Int, dim::Int)
  cols = [m for m in 1:size(terms, 2) if Int(terms[1, m]) == coord]
  function q(t, args...)
    ξ = args[1:dim]                      # (t, ξ[, η[, ζ]], p)
    acc = zero(promote_type(map(typeof, ξ)...))
    for m in cols
      c = terms[2, m] + terms[3, m] * t + terms[4, m] * sin(terms[5, m] * t + terms[6, m])
      v = c * one(acc)
      for d in 1:dim
        v *= factor(Int(terms[7 + 3 * (d - 1), m]), terms[8 + 3 * (d - 1), m], terms[9 + 3 * (d - 1), m], ξ[d])
      end
      acc += v
    end
    return acc
  end
  return q
end

# the NON-separable test mappings (same formulas: oracle/metric_oracle.hpp FuncMap ids 1-3, workloads.GENERIC_TEST_SOURCES);
# p is the parameter vector, passed as the grid's `params`
function funcmap_closures(id::Int)
  if id == 1
    x = (t, ξ, η, ζ, p) -> ξ + p[1] * sin(p[2] * (η + 0.5 * ζ) + 0.3 * t)
    y = (t, ξ, η, ζ, p) -> η * (1.0 + p[3] * ξ) / (1.0 + p[4] * ζ * ζ)
    z = (t, ξ, η, ζ, p) -> ζ + p[1] * exp(-p[5] * (ξ - η) * (ξ - η)) * cos(t) + sqrt(1.0 + 0.01 * ξ * ξ)
    return (x, y, z)
  elseif id == 2
    x = (t, ξ, η, p) -> ξ + p[1] * sin(p[2] * ξ * η + t)
    y = (t, ξ, η, p) -> η + p[3] * sqrt(1.0 + p[4] * ξ * ξ + 0.5 * η * η)
    return (x, y)
  else
    return ((t, ξ, p) -> ξ + p[1] * sin(p[2] * ξ) * exp(-p[3] * ξ) + p[4] * t,)
  end
end

dump(path, A) = write(path, collect(reinterpret(Float64, vec(Array(A)))))

function dump_grid(dir, grid, dim)
  mkpath(dir)
  cm = cell_metrics(grid)
  fm = face_metrics(grid)
  dump(joinpath(dir, "cell_forward.bin"), cm.forward)
  dump(joinpath(dir, "cell_inverse.bin"), cm.inverse)
  for a in 1:dim
    dump(joinpath(dir, "face$(a - 1)_forward.bin"), fm[a].forward)
    dump(joinpath(dir, "face$(a - 1)_inverse.bin"), fm[a].inverse)
    dump(joinpath(dir, "face$(a - 1)_conserved.bin"), fm[a].conserved)
  end
  for c in 1:dim
    dump(joinpath(dir, "node$(c - 1).bin"), grid.node_coordinates[c])
    dump(joinpath(dir, "centroid$(c - 1).bin"), grid.centroid_coordinates[c])
    for a in 1:dim
      dump(joinpath(dir, "face$(a - 1)coord$(c - 1).bin"), grid.face_coordinates[a][c])
    end
  end
  open(joinpath(dir, "done.txt"), "w") do f
    println(f, "CurvilinearGrids ", pkgversion(CurvilinearGrids), " julia ", VERSION)
  end
end

for name in sort(readdir(INP))
  dir = joinpath(INP, name)
  isdir(dir) || continue
  meta = read_case(dir)
  dim, nhalo = parse(Int, meta["dim"]), parse(Int, meta["nhalo"])
  for sname in split(meta["schemes"], ",")
    scheme = scheme_of(sname)
    out = joinpath(OUT, name, sname)
    if meta["kind"] == "discrete"
      nn = ints(meta["nn"])
      nodes = [readf64(joinpath(dir, "xyz"[q:q] * ".bin"), nn...) for q in 1:dim]
      kw = (; halo_coords_included=meta["halo_coords_included"] == "1", cache_mode=:eager,
            conserved_metric_scheme=scheme, coordinate_system=cs_of(meta["coordinate_system"]),
            basis=basis_of(meta["basis"]))
      grid = dim == 1 ? DiscreteGrid(vec(nodes[1]), nhalo; kw...) : DiscreteGrid(nodes..., nhalo; kw...)
      dump_grid(out, grid, dim)
    elseif meta["kind"] == "funcmap"
      celldims = ints(meta["celldims"])
      p = parse.(Float64, split(meta["params"], ","))
      t = parse(Float64, meta["t"])
      grid = MappedGrid(funcmap_closures(parse(Int, meta["map_id"]))..., p, celldims, nhalo; cache_mode=:lazy,
                        conserved_metric_scheme=scheme, t=0.0)
      update!(grid, t, p)
      dump_grid(out, grid, dim)
    else
      celldims = ints(meta["celldims"])
      terms = readf64(joinpath(dir, "terms.bin"), 15, parse(Int, meta["nterms"]))
      t = parse(Float64, meta["t"])
      maps = [term_mapping(terms, c - 1, dim) for c in 1:dim]
      # the time-dependent case goes through update! like BASELINE config 4 (mapped_grid.jl:551-593)
      grid = MappedGrid(maps..., (;), celldims, nhalo; cache_mode=:lazy, conserved_metric_scheme=scheme, t=0.0)
      update!(grid, t, (;))
      dump_grid(out, grid, dim)
    end
    println("wrote ", out)
  end
end
