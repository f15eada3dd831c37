"""mlfortran_b200 -- B200-native drop-in for MlFortran's Gaussian-mixture EM step object."""
from .emgmm import mlf_algo_emgmm, mlf_evalGaussian, mlf_hdf5_file, mlf_init, mlf_maxGaussian, mlf_quit  # noqa: F401
