echo OLD; python tools/_fit_ab.py /root/repo/_old 2>&1 | tail -4
echo NEW; python tools/_fit_ab.py /root/repo 2>&1 | tail -4
