# instruction shares per code region of the summary kernel (line ranges of the sources at the commit that took the capture)
python scripts/ncu_regions2.py "$1" \
  fo_metric_dev.cuh:187-219=erfc fo_metric_dev.cuh:221-256=cp_gauss_terms fo_metric_dev.cuh:293-343=be_prepare \
  fo_metric_dev.cuh:345-349=be_arclen fo_metric_dev.cuh:350-450=be_bisect fo_metric_dev.cuh:167-175=obb_hit \
  fo_metric_dev.cuh:61-92=obb_d2 fo_metric_dev.cuh:257-272=logistic+lr4s_coef \
  fo_metric_sweep.cu:136-141=lr4s_atan2 fo_metric_sweep.cu:143-158=harm_logits fo_metric_sweep.cu:222-257=ego_staging+windows \
  fo_metric_sweep.cu:279-297=bound_updates fo_metric_sweep.cu:298-355=drain_near fo_metric_sweep.cu:356-393=drain_cp \
  fo_metric_sweep.cu:394-481=per_step_loop fo_metric_sweep.cu:482-563=window_filter fo_metric_sweep.cu:564-585=pool \
  fo_metric_sweep.cu:586-642=tile_epilogue+BE_list fo_metric_sweep.cu:643-701=traj_epilogue
