"""xpsi_b200 -- B200-native likelihood hot path behind the X-PSI call signatures.

Sub-modules mirror the reference layout for the path (same callables, same
positional order, same return conventions):

    xpsi_b200.cellmesh.integrator_for_azimuthal_invariance.integrate
    xpsi_b200.cellmesh.integrator.integrate
    xpsi_b200.cellmesh.integrator_for_time_invariance.integrate
    xpsi_b200.surface_radiation_field.intensity
    xpsi_b200.tools.energy_integrator / energy_interpolator / phase_integrator / phase_interpolator
    xpsi_b200.tools.synthesise_exposure / synthesise_given_total_count_number
    xpsi_b200.instrument.fold / Instrument
    xpsi_b200.interstellar.Interstellar
    xpsi_b200.likelihoods.precomputation / eval_marginal_likelihood / poisson_likelihood_given_background
    xpsi_b200.likelihood.Likelihood                 (xpsi.Likelihood.__call__ over the pipeline)
    xpsi_b200.pipeline.BatchedLikelihood            (additional, batched; embed on the GPU; several signals, Elsewhere,
                                                     Everywhere, interstellar attenuation)
    xpsi_b200.from_xpsi.from_xpsi                   (GPU likelihood from a constructed xpsi.Likelihood object)
    xpsi_b200.dropin.install                        (rebinds the compiled seams inside an imported xpsi package)
    xpsi_b200.sampling                              (vectorised prior, UltraNest-shaped callables, sharded sweep, importance)

Everything computes on the GPU through libxpsi_b200.so; there is no CPU path.
``xpsi_b200.synthetic`` (pure numpy) defines the synthetic workloads.
"""
__version__ = "0.1.0"
