"""Discovery writer - drop-in for rc_frontend/redis_channel_publisher.py:20-93.

Every second the reference frontend does, in one pipeline (rc_frontend/redis_channel_publisher.py:84-91):

    SADD channelizers <instance_uuid>
    SET  <instance_uuid> json{instance_uuid, start_time, current_time, hostname, pid, address, port,
                              channel_count, source_count, sources: [[center_freq, samp_rate], ...], index?}

which is how `frontend_connector` instances find a frontend's REP port.  Same constructor keywords and the same
background thread here; the only additions are an injectable `client` (anything with `.pipeline()` returning an
object with `.sadd / .set / .execute`, e.g. redis.StrictRedis or the in-process fake of the tests - the `redis`
package is an optional dependency of this frontend) and `publish_once()` so the loop body can be driven
synchronously.  Errors while publishing are logged, never raised (reference :90-93).
"""
import json
import logging
import os
import socket
import threading
import time
import uuid
from urllib.parse import urlparse


class redis_channel_publisher(object):
    def __init__(self, host=None, port=None, sources=None, channels=None, zmq_socket=None, index=None,
                 client=None, instance_uuid=None, interval=1.0, start=True):
        self.log = logging.getLogger("redis_channel_publisher")
        self.host = host if host is not None else "127.0.0.1"
        self.port = port if port is not None else 6379
        if sources is None:
            raise Exception("Sources must be provided at initialization")
        if channels is None:
            raise Exception("Channels must be provided at initialization")
        if zmq_socket is None:
            raise Exception("ZMQ Socket must be provided at initialization")
        self.start_time = time.time()
        self.sources = sources
        self.channels = channels
        self.zmq_socket = zmq_socket
        self.index = index
        self.instance_uuid = instance_uuid or str(uuid.uuid4())
        self.interval = interval
        self.client = client
        self.continue_running = True
        self.published = 0
        if self.client is None:
            self.init_connection()
        self._thread = None
        if start:
            self._thread = threading.Thread(target=self.publish_loop, name="redis-publisher", daemon=True)
            self._thread.start()

    def init_connection(self):
        import redis  # optional dependency, exactly what the reference imports
        self.client = redis.StrictRedis(host=self.host, port=self.port, db=0)

    def _rep_port(self):
        import zmq
        ep = self.zmq_socket.getsockopt(zmq.LAST_ENDPOINT)
        if isinstance(ep, bytes):
            ep = ep.decode("utf-8")
        return urlparse(ep).port

    def _address(self):
        try:
            return socket.gethostbyname(socket.gethostname())
        except OSError:   # container hostnames do not always resolve; the reference would raise here
            return "127.0.0.1"

    def publish_data(self):
        data = {
            "instance_uuid": self.instance_uuid,
            "start_time": self.start_time,
            "current_time": time.time(),
            "hostname": socket.gethostname(),
            "pid": os.getpid(),
            "address": self._address(),
            "port": self._rep_port(),
            "channel_count": len(self.channels),
            "source_count": len(self.sources),
            "sources": [],
        }
        if self.index is not None:
            data["index"] = self.index
        for source_id in self.sources:
            source = self.sources[source_id]
            data["sources"].append((source["center_freq"], source["samp_rate"]))
        return data

    def publish_once(self):
        pipe = self.client.pipeline()
        pipe.sadd("channelizers", self.instance_uuid)
        pipe.set(self.instance_uuid, json.dumps(self.publish_data()))
        # (expiry is handled by the manager side, reference :87-88)
        try:
            pipe.execute()
            self.published += 1
            return True
        except Exception as e:
            self.log.error("Exception submitting redis demod publish: %s" % e)
            return False

    def publish_loop(self):
        self.log.info("publish_loop() startup")
        time.sleep(min(0.5, self.interval))
        while self.continue_running:
            self.publish_once()
            time.sleep(self.interval)

    def stop(self):
        self.continue_running = False
        if self._thread is not None:
            self._thread.join(timeout=2.0)
