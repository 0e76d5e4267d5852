cd /root/repo
python profiles/scripts/probe_fc_csr.py csr
python profiles/scripts/probe_fc_csr.py fc5
