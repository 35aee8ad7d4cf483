TAG=r02_final_cfg2 WORKLOAD=cfg2 KREGEX='hupd_ts_kernel|recon_ts_kernel|gradw_ts_kernel' bash tools/gpu_profile.sh
TAG=r02_final_cfg3 WORKLOAD=cfg3 KREGEX='hupd_ts_kernel|recon_os_kernel|gradw_ns_kernel' bash tools/gpu_profile.sh
