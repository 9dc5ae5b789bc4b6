"""Model construction exactly as the reference drivers do it (scripts/run_ala.py:184-209, scripts/run_pdb.py:330-333)."""
from torch import nn

from .cgvae import CGequiVAE, CGprior, EquiEncoder, EquivariantDecoder, EquivariantPsuedoDecoder, PCN


def build_cgvae(n_basis, n_rbf, enc_nconv, dec_nconv, atom_cutoff, cg_cutoff, n_cgs, activation="swish", det=False):
    """decoder gets cutoff=atom_cutoff, encoder / prior get cutoff=cg_cutoff; breaksym iff n_cgs == 3 (run_ala.py:192-206)."""
    atom_mu = nn.Sequential(nn.Linear(n_basis, n_basis), nn.ReLU(), nn.Linear(n_basis, n_basis))
    atom_sigma = nn.Sequential(nn.Linear(n_basis, n_basis), nn.ReLU(), nn.Linear(n_basis, n_basis))
    decoder = EquivariantPsuedoDecoder(n_atom_basis=n_basis, n_rbf=n_rbf, cutoff=atom_cutoff, num_conv=dec_nconv,
                                       activation=activation, breaksym=(n_cgs == 3))
    encoder = EquiEncoder(n_conv=enc_nconv, n_atom_basis=n_basis, n_rbf=n_rbf, cutoff=cg_cutoff, activation=activation,
                          cg_mp=False, dir_mp=False)
    prior = CGprior(n_conv=enc_nconv, n_atom_basis=n_basis, n_rbf=n_rbf, cutoff=cg_cutoff, activation=activation, dir_mp=False)
    return CGequiVAE(encoder, decoder, atom_mu, atom_sigma, n_cgs, feature_dim=n_basis, prior_net=prior, det=det,
                     equivariant=True)


def build_pcn(n_basis, n_rbf, cg_cutoff, dec_nconv, activation="swish", cross=True):
    """run_pdb.py:330-333"""
    decoder = EquivariantDecoder(n_atom_basis=n_basis, n_rbf=n_rbf, cutoff=cg_cutoff, num_conv=dec_nconv,
                                 activation=activation, cross_flag=cross)
    return PCN(decoder, feature_dim=n_basis, offset=False)
