"""Attribute-style configs equal to the three in-scope reference YAML files
(config/qm8_lanczos_net.yaml, config/qm8_ada_lanczos_net.yaml, config/graph_lanczos_net.yaml);
only the fields the model constructors read (model/lanczos_net.py:18-35 etc.)."""
from types import SimpleNamespace as NS


def qm8_lanczos_net(**model_over):
  model = dict(name='LanczosNet', short_diffusion_dist=[],
               long_diffusion_dist=[1, 2, 3, 5, 7, 10, 20, 30], num_eig_vec=20,
               spectral_filter_kind='MLP', input_dim=64, hidden_dim=[128] * 7, output_dim=16,
               num_layer=7, loss='MSE', output_func='MLP')
  model.update(model_over)
  return NS(seed=1234, dataset=NS(loader_name='QM8Data', name='chemistry', num_atom=70,
                                  num_bond_type=6), model=NS(**model))


def qm8_gcn(**model_over):
  """config/qm8_gcn.yaml"""
  model = dict(name='GCN', input_dim=64, hidden_dim=[128] * 7, output_dim=16, num_layer=7,
               loss='MSE', output_func='MLP')
  model.update(model_over)
  return NS(seed=1234, dataset=NS(loader_name='QM8Data', name='chemistry', num_atom=70,
                                  num_bond_type=6), model=NS(**model))


def qm8_dcnn(**model_over):
  """config/qm8_dcnn.yaml"""
  model = dict(name='DCNN', input_dim=64, diffusion_dist=[3, 5, 7, 10, 20, 30], hidden_dim=[128] * 7,
               output_dim=16, num_layer=7, loss='MSE', output_func='MLP')
  model.update(model_over)
  return NS(seed=1234, dataset=NS(loader_name='QM8Data', name='chemistry', num_atom=70,
                                  num_bond_type=6), model=NS(**model))


def qm8_cheby_net(**model_over):
  """config/qm8_cheby_net.yaml"""
  model = dict(name='ChebyNet', input_dim=64, polynomial_order=5, hidden_dim=[128] * 7,
               output_dim=16, num_layer=7, loss='MSE', output_func='MLP')
  model.update(model_over)
  return NS(seed=1234, dataset=NS(loader_name='QM8Data', name='chemistry', num_atom=70,
                                  num_bond_type=6), model=NS(**model))


def qm8_ada_lanczos_net(**model_over):
  model = dict(name='AdaLanczosNet', short_diffusion_dist=[1, 2, 3],
               long_diffusion_dist=[5, 7, 10, 20, 30], num_eig_vec=20,
               use_reorthogonalization=False, use_power_iteration_cap=False,
               spectral_filter_kind='MLP', input_dim=64, hidden_dim=[128] * 7, output_dim=16,
               num_layer=7, loss='MSE', output_func='MLP')
  model.update(model_over)
  return NS(seed=1234, dataset=NS(loader_name='QM8Data', name='chemistry', num_atom=70,
                                  num_bond_type=6), model=NS(**model))


def graph_lanczos_net(**model_over):
  model = dict(name='LanczosNetGeneral', short_diffusion_dist=[],
               long_diffusion_dist=[1, 2, 3, 5, 7, 10, 20, 30], num_eig_vec=20,
               spectral_filter_kind='MLP', input_dim=10, hidden_dim=[128] * 7, output_dim=2,
               num_layer=7, loss='MSE', output_func='MLP')
  model.update(model_over)
  return NS(seed=1234, dataset=NS(loader_name='GraphData', name='synthetic', node_emb_dim=10,
                                  graph_emb_dim=2, num_edge_type=1), model=NS(**model))
