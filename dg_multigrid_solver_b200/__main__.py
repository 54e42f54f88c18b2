"""`python -m dg_multigrid_solver_b200 <flags>` -- same command line as `python -m dgfem`
(dgfem/__main__.py:6-74): run it from a directory that holds input/paramfile.yml and the grid
folder the paramfile names.  -m (multigrid) and -s --smoother X run on the B200; the other
solver flags parse but report that they are outside the accelerated path."""
import argparse
import logging
import sys
import traceback


class MutuallyInclusiveArgumentError(Exception):
    pass


class MutuallyExclusiveArgumentError(Exception):
    pass


def build_parser():
    parser = argparse.ArgumentParser(prog="DG solver", description="DG solver for the Poisson and Stokes problems")
    parser.add_argument("--grid-folder", type=str)
    parser.add_argument("-f", "--grid-file", type=str)
    parser.add_argument("--p-grid", type=int)
    parser.add_argument("--p-solution", type=int)
    solver = parser.add_mutually_exclusive_group(required=True)
    solver.add_argument("-d", "--solve-direct", action="store_true")
    solver.add_argument("-s", "--solve-smoother", help="mutually inclusive with --smoother", action="store_true")
    parser.add_argument("--smoother", type=str)
    solver.add_argument("-amg", "--solve-pyamg", action="store_true")
    solver.add_argument("-k", "--solve-krylov", action="store_true")
    solver.add_argument("-m", "--solve-multigrid", action="store_true")
    solver.add_argument("-fvm", "--solve-finite-volume-method", action="store_true")
    solver.add_argument("-amp", "--solve-smoother-amplification", action="store_true")
    parser.add_argument("--dg-discretization", action="store_true")
    parser.add_argument("--fvm-discretization", action="store_true")
    parser.add_argument("--check-eigenvalues", action="store_true")
    parser.add_argument("--check-condition-number", action="store_true")
    parser.add_argument("--plot-sparsity-pattern", action="store_true")
    parser.add_argument("-v", "--verbose", action="store_true")
    parser.add_argument("--silent", action="store_true")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.solve_smoother and not args.smoother:
        raise MutuallyInclusiveArgumentError("--solve-smoother option must be used with --smoother")
    from .settings import Settings, load_params
    settings = Settings(load_params())
    if args.verbose:
        settings.update_setting("logging.loglevel", "DEBUG")
    if args.silent:
        settings.update_setting("logging.loglevel", "ERROR")
    logging.basicConfig(level=getattr(logging, settings.logging.loglevel, logging.INFO),
                        format="%(asctime)s %(name)s[%(process)d] %(levelname)s %(message)s")
    logger = logging.getLogger("dgfem")
    logger.info("starting DG-FEM")
    from .dgfem import DGFEM
    try:
        d = DGFEM(settings=settings, grid_folder=args.grid_folder, grid_file=args.grid_file, p_grid=args.p_grid,
                  p_solution=args.p_solution, solve_direct=args.solve_direct, solve_smoother=args.solve_smoother,
                  solve_smoother_amplification=args.solve_smoother_amplification, solve_pyamg=args.solve_pyamg,
                  solve_krylov=args.solve_krylov, solve_multigrid=args.solve_multigrid,
                  solve_finite_volume_method=args.solve_finite_volume_method, smoother=args.smoother,
                  check_eigenvalues=args.check_eigenvalues, check_condition_number=args.check_condition_number,
                  plot_sparsity_pattern=args.plot_sparsity_pattern)
        d.solve()
        logger.info(f"Solving with {settings.solver.method} method took {d.solver.timings.get('solve', 0.0):.4g} seconds")
        logger.info(f"L2 norm of the residual (modal): {d.residual:.6e} (not normalized)")
        logger.info(f"L2 norm of the residual (modal): {d.residual_normalized:.6e} (normalized)")
        if hasattr(d, "L1_error_u"):
            logger.info(f"The norms of the error (nodal) are: L1={d.L1_error_u:.6e}, L2={d.L2_error_u:.6e}")
        if settings.solver.method == "multigrid":
            logger.info(f"V-cycles: {len(d.solver.residuals) - 1}")
        return d
    except Exception:
        logger.critical(traceback.format_exc())        # dgfem/__main__.py:73-74
        return None


if __name__ == "__main__":
    main(sys.argv[1:])
